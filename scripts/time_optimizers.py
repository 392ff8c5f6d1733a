"""Time the device-side design updates (OC bisection, MMA subproblem) at the size of BASELINE configs[2]
(4 194 304 design variables, one volume constraint) on synthetic sensitivities.  Diagnostic, not part of bench.py's
metric:  python scripts/time_optimizers.py [nel]  -> one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import __graft_entry__ as ge

    ge.build()
    import pymoto_b200 as pmb
    from pymoto_b200 import device as dv
    from pymoto_b200.optimizers import MmaDeviceOps, mma_design_update

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256 * 128 * 128
    rng = np.random.default_rng(7)
    x = dv.to_device(np.clip(0.5 + 0.2 * (rng.random(n) - 0.5), 0, 1))
    dg0 = dv.to_device(-np.abs(rng.standard_normal(n)) * 1e-3 - 1e-5)   # compliance-like: negative
    dg1 = dv.to_device(np.full(n, 10.0 / (0.5 * n)))                     # scaled volume constraint
    g = np.array([100.0, -0.01])
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    # ---- MMA (MMA2007, m = 1): three successive updates so the asymptote pass is included
    ops = MmaDeviceOps(n, 1)
    offset = dv.to_device(np.full(n, 0.5))
    opt = dict(albefa=0.1, asyincr=1.2, asydecr=0.7, asybound=10.0, a0=1.0, epsimin=1e-10, rho=1e-5, version=2007,
               a=np.zeros(1), c=np.full(1, 1e3), d=np.ones(1))
    xold1 = xold2 = None
    mma_ms, its = [], []
    for k in range(4):
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        t0 = time.perf_counter()
        e0.record()
        lam, nit = mma_design_update(ops, x, g, [dg0, dg1], offset, xold1, xold2, 0.0, 1.0, 0.1, opt)
        e1.record()
        torch.cuda.synchronize()
        mma_ms.append((e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)))
        its.append(nit)
        xold2, xold1 = xold1, x.clone()
        x = ops.x.clone()
    # ---- OC
    oc = pmb.OC(pmb.Signal("x", state=x), pmb.Signal("c", state=1.0), pmb.Network(), verbosity=0)
    oc_ms = []
    for k in range(3):
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        xn = oc._update(x, dg0)
        e1.record()
        torch.cuda.synchronize()
        oc_ms.append(e0.elapsed_time(e1))
    print(json.dumps({"n": n, "mma_update_ms_device_wall": mma_ms[1:], "mma_newton_iterations": its[1:],
                      "oc_update_ms": oc_ms[1:], "note": "MMA2007, 1 constraint; each Newton iteration = 3+ fused passes over 17 vectors"}))


if __name__ == "__main__":
    main()
