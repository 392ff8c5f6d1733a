PMB_DEBUG=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dense_inverse" 2>&1 | grep -E "pmb_dense|passed|failed" | head -5
cat > /tmp/t.py <<'PY'
import sys, torch, numpy as np, ctypes as C
sys.path.insert(0,'.')
import pymoto_b200 as pmb
from pymoto_b200 import _lib, device as dv
n=675
rng=np.random.default_rng(0)
M=rng.standard_normal((n,n)); M=M@M.T+n*np.eye(n)
for rep in range(2):
    A=dv.to_device(M.copy().ravel()); scr=dv.empty(_lib.query("pmb_dense_invert_ws_doubles", n)); info=dv.zeros(1, torch.int32)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); _lib.call("pmb_dense_invert", n, dv.ptr(A), dv.ptr(scr), dv.ptr(info), dv.stream()); e1.record(); torch.cuda.synchronize()
    print("dense_invert ms", e0.elapsed_time(e1), "err", np.abs(A.cpu().numpy().reshape(n,n)@M-np.eye(n)).max())
PY
PMB_DEBUG=1 python /tmp/t.py; PMB_DENSE_COOP=0 python /tmp/t.py
