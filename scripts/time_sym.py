#!/usr/bin/env python
"""Time the level-1 Jacobi sweep of the bench hierarchy in both layouts (stencil-CSR tile kernel / symmetric half-stencil).

  python scripts/time_sym.py [--out gpurun_out/time_sym.json] [--grid 256x128x128] [--ndof 3]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--grid", default="256x128x128")
    ap.add_argument("--ndof", type=int, default=3)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    import __graft_entry__ as ge

    ge.build()
    import pymoto_b200 as pmb
    from pymoto_b200 import _lib, device as dv
    from pymoto_b200.matrix import DeviceCSR

    os.environ["PMB_ELEM_AUTOTUNE"] = "0"
    nx, ny, nz = (int(v) for v in args.grid.split("x"))
    ndof = args.ndof
    dom = pmb.VoxelDomain(nx, ny, nz)
    nodes_face = (np.arange(nz + 1)[:, None] * (ny + 1) + np.arange(ny + 1)[None, :]).ravel() * (nx + 1)
    bc = np.sort((nodes_face[:, None] * ndof + np.arange(ndof)[None, :]).ravel())
    asm = (pmb.AssembleStiffness if ndof == 3 else pmb.AssemblePoisson)(dom, bc=bc)
    x = torch.rand(dom.nel, dtype=torch.float64, device="cuda") * 0.9 + 0.1
    K = asm(x)
    DeviceCSR.symmetric_storage = True  # (off by default: build the symmetric copy for this measurement)
    mg = pmb.solvers.auto_multigrid(dom)[0]
    mg.update(K)
    A1 = mg.Ac
    n = A1.shape[0]
    D = A1.diagonal_device()
    u, u2, b = A1.new_vec(zero=True), A1.new_vec(zero=True), A1.new_vec(zero=True)
    u.copy_(torch.rand(n, dtype=torch.float64, device="cuda"))
    b.copy_(torch.rand(n, dtype=torch.float64, device="cuda"))
    res = {"grid": args.grid, "ndof": ndof, "level1_rows": n, "nnz": A1.nnz, "asymmetry": getattr(A1, "asymmetry", None),
           "sym_valid": A1._sym_valid}
    ref = None
    for name, on in (("stencil_csr", False), ("symmetric", True)):
        DeviceCSR.symmetric_storage = on
        for _ in range(3):
            A1.apply(_lib.JACOBI, u, u2, b=b, diag=D, w=0.5)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            A1.apply(_lib.JACOBI, u, u2, b=b, diag=D, w=0.5)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        y = u2.clone()
        if ref is None:
            ref = y
        nb = 27 if not on else 14
        bytes_ = 8 * (A1.nnz * nb / 27) + 32 * n
        res[name] = {"ms": ms, "algorithmic_bytes": bytes_, "gbs": bytes_ / (ms * 1e-3) / 1e9,
                     "maxdiff_vs_stencil_csr": float((y - ref).abs().max())}
    txt = json.dumps(res, indent=1)
    print(txt)
    if args.out:
        open(args.out, "w").write(txt)


if __name__ == "__main__":
    main()
