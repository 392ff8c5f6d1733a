#!/usr/bin/env python
"""Time every layout of the matrix-free finest-level kernel (pmb_elem_spmv, Jacobi mode) at bench sizes.

  python scripts/time_elem.py [--out gpurun_out/time_elem.json]
Prints one JSON object: per (grid, ndof) the average launch time of each layout (CUDA events, 20 launches after 3 warm-up,
operands larger than L2 are not the point here: the kernel is compute / latency bound) and the FP64 rate it implies.
"""
import argparse
import ctypes as C
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
    ap.add_argument("--cases", default="3:256x128x128,1:256x256x256")
    ap.add_argument("--variants", default=None, help="comma-separated layouts (default: all)")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--nobc", action="store_true", help="no Dirichlet dofs (no mask traffic at all): the kernels' upper bound")
    args = ap.parse_args()
    import __graft_entry__ as ge

    ge.build()
    import pymoto_b200 as pmb
    from pymoto_b200 import _lib, device as dv

    os.environ["PMB_ELEM_AUTOTUNE"] = "0"
    res = {}
    for case in args.cases.split(","):
        ndof, dims = case.split(":")
        ndof = int(ndof)
        nx, ny, nz = (int(v) for v in dims.split("x"))
        dom = pmb.VoxelDomain(nx, ny, nz)
        nodes_face = (np.arange(nz + 1)[:, None] * (ny + 1) + np.arange(ny + 1)[None, :]).ravel() * (nx + 1)
        bc = np.sort((nodes_face[:, None] * ndof + np.arange(ndof)[None, :]).ravel())
        asm = (pmb.AssembleStiffness if ndof == 3 else pmb.AssemblePoisson)(dom, bc=None if args.nobc else bc)
        x = torch.rand(dom.nel, dtype=torch.float64, device="cuda") * 0.9 + 0.1
        K = asm(x)
        gen = K.generator
        n = K.shape[0]
        D = K.diagonal_device()
        u, u2, b = K.new_vec(zero=True), K.new_vec(zero=True), K.new_vec(zero=True)
        u.copy_(torch.rand(n, dtype=torch.float64, device="cuda"))
        b.copy_(torch.rand(n, dtype=torch.float64, device="cuda"))
        out = {}
        nvar = _lib.query("pmb_elem_num_variants")
        ref = None
        for v in (range(nvar) if args.variants is None else [int(t) for t in args.variants.split(",")]):
            gen.variant = v
            for _ in range(3):
                K.apply(_lib.JACOBI, u, u2, b=b, diag=D, w=0.5)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                K.apply(_lib.JACOBI, u, u2, b=b, diag=D, w=0.5)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            y = u2.clone()
            same = None if ref is None else bool(torch.equal(y, ref))
            if ref is None:
                ref = y
            maxdiff = float((y - ref).abs().max())
            flops = 2.0 * 8 * (8 * ndof * ndof + ndof) * (n / ndof)
            out[f"variant{v}"] = {"ms": ms, "fp64_tflops": flops / (ms * 1e-3) / 1e12, "bit_identical_to_0": same, "maxdiff": maxdiff,
                                  "hbm_gbs_algorithmic": (40 * n + 8 * dom.nel) / (ms * 1e-3) / 1e9}
        res[case] = out
        del K, asm, u, u2, b, D, x
        torch.cuda.empty_cache()
    txt = json.dumps(res, indent=1)
    print(txt)
    if args.out:
        open(args.out, "w").write(txt)


if __name__ == "__main__":
    main()
