"""Device plumbing: torch owns device memory and streams; libpmb owns the arithmetic.

No CPU fallback exists anywhere in this package: constructing any device object without CUDA raises.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.PmbError("pymoto_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    _lib.load()
    return torch.device("cuda", torch.cuda.current_device())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Raw device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def is_device(a):
    return isinstance(a, torch.Tensor) and a.is_cuda


def to_device(a, dtype=torch.float64):
    """numpy array / tensor -> contiguous CUDA tensor (copies host arrays; passes device tensors through)."""
    dev = require_cuda()
    if isinstance(a, torch.Tensor):
        t = a.to(device=dev, dtype=dtype)
    else:
        arr = np.ascontiguousarray(np.asarray(a))
        if np.iscomplexobj(arr):
            raise TypeError("complex values are not supported by the B200 hot path (real FP64 only)")
        t = torch.from_numpy(arr).to(device=dev, dtype=dtype)
    return t.contiguous()


def like_input(result, template):
    """Return ``result`` (a CUDA tensor) in the residency of ``template``: numpy in -> numpy out."""
    if is_device(template):
        return result
    return result.cpu().numpy()


def empty(n, dtype=torch.float64):
    return torch.empty(int(n), dtype=dtype, device=require_cuda())


def zeros(n, dtype=torch.float64):
    return torch.zeros(int(n), dtype=dtype, device=require_cuda())


class _Workspace:
    """Per-device scratch: reduction workspace (zero-initialised once, self re-arming) and a scalar arena."""

    def __init__(self):
        self.dev = require_cuda()
        self.red = torch.zeros(_lib.query("pmb_ws_doubles"), dtype=torch.float64, device=self.dev)
        self._spmv_ws = None
        self._gal_ws = None

    def spmv_ws(self, n):
        if self._spmv_ws is None or self._spmv_ws.numel() < n:
            self._spmv_ws = torch.empty(int(n), dtype=torch.float64, device=self.dev)
        return self._spmv_ws

    def galerkin_ws(self, n):
        """Scratch for the two-pass Galerkin product (shared by all levels: the finest level sizes it)."""
        if self._gal_ws is None or self._gal_ws.numel() < n:
            self._gal_ws = None
            self._gal_ws = torch.empty(int(n), dtype=torch.float64, device=self.dev)
        return self._gal_ws


_workspaces = {}


def workspace():
    dev = require_cuda()
    ws = _workspaces.get(dev.index)
    if ws is None:
        ws = _workspaces[dev.index] = _Workspace()
    return ws


# ------------------------------------------------------------------------------------------------ vector ops
def dots(pairs):
    """Deterministic device dot products of up to 4 (a, b) pairs; returns a fresh device tensor of len(pairs)."""
    k = len(pairs)
    out = empty(4)
    n = pairs[0][0].numel()
    args = []
    for i in range(4):
        if i < k:
            args += [ptr(pairs[i][0]), ptr(pairs[i][1])]
        else:
            args += [None, None]
    _lib.call("pmb_dots", n, k, *args, ptr(out), ptr(workspace().red), stream())
    return out[:k]


def lincomb(out, ca, a, cb=None, b=None):
    """out = ca*a + cb*b with coefficients that may reference device scalars (see _lib.coef)."""
    if not isinstance(ca, _lib.Coef):
        ca = _lib.coef(ca)
    if cb is None:
        cb = _lib.coef(0.0)
    elif not isinstance(cb, _lib.Coef):
        cb = _lib.coef(cb)
    _lib.call("pmb_lincomb", out.numel(), ptr(out), ca, ptr(a), cb, ptr(b), stream())
    return out


def scalar_ptr(t, i=0):
    """Device address of element i of a float64 tensor, for _lib.coef(num=..., den=...)."""
    return t.data_ptr() + 8 * i


def sm_count():
    """Streaming multiprocessors of the current device (148 on a B200)."""
    require_cuda()
    return torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
