"""Probe: latency of one halo exchange / one small all-reduce on the z-slab path, per transport:
one-launch peer kernels (pmb_peer_*), three-launch mailboxes (copy + symmetric-memory barrier + copy), NCCL.
Launched eagerly and replayed from a CUDA graph (how the V-cycle issues them).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29530 scripts/probe_peer_xchg.py
"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
import pymoto_b200 as pmb  # noqa: E402

dom = pmb.VoxelDomain(256, 128, 32 * world)
REP = 50


def timed(fn, n):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n, 1e6 * (time.perf_counter() - t0) / n  # device us, wall us


for label, env in [("one-launch", {"PMB_PEER_FUSED": "1"}), ("three-launch", {"PMB_PEER_FUSED": "0"}), ("nccl", {"PMB_HALO_MAILBOX": "0"})]:
    os.environ.pop("PMB_PEER_FUSED", None); os.environ.pop("PMB_HALO_MAILBOX", None)
    os.environ.update(env)
    ctx = pmb.slab.init(dom, n_levels=1)
    comm = ctx.comm
    for lvl, plane in enumerate([257 * 129 * 3, 129 * 65 * 3, 65 * 33 * 3]):
        own = 4 * plane
        buf = torch.zeros(own + 2 * plane, dtype=torch.float64, device=dev)
        buf[plane:plane + own] = rank + 1.0
        ex = lambda: comm.exchange(buf, plane, own, plane)  # noqa: E731
        dev_us, wall_us = timed(ex, 200)
        line = f"{label:12s} level-{lvl} plane ({8 * plane / 1e3:.0f} kB): eager {dev_us:6.1f} us (host issue {wall_us:6.1f} us)"
        if label != "nccl":
            torch.cuda.synchronize(); dist.barrier()
            g = torch.cuda.CUDAGraph()
            comm.begin_capture()
            with torch.cuda.graph(g):
                for _ in range(REP):
                    ex()
                comm.end_capture()
            d2, _ = timed(g.replay, 8)
            line += f", graph-replayed {d2 / REP:6.1f} us"
        if rank == 0:
            print(line, flush=True)
    v = torch.ones(4, dtype=torch.float64, device=dev)
    d, w = timed(lambda: comm.allreduce_(v.clone()), 200)
    if rank == 0:
        print(f"{label:12s} all-reduce of 4 doubles: {d:6.1f} us (host issue {w:6.1f} us), fast path used: {comm.fast_allreduces > 0}", flush=True)
    comm.check_peer_timeouts()
    pmb.slab.reset()
dist.barrier()
dist.destroy_process_group()
