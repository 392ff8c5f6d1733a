"""Probe: halo exchange through symmetric memory (peer copies + device barrier) vs NCCL send/recv batches.
torchrun --nproc-per-node 2 scripts/probe_symm_halo.py"""
import os
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
lower = rank - 1 if rank > 0 else None
upper = rank + 1 if rank < world - 1 else None

for plane in [25_000, 100_000, 400_000]:
    n = plane * 8
    base = torch.zeros(n + 2 * plane, dtype=torch.float64, device=dev)
    base[plane:plane + n] = rank + torch.arange(n, device=dev, dtype=torch.float64) * 1e-9
    # ---- NCCL reference
    def nccl_exchange():
        ops = []
        if lower is not None:
            ops += [dist.P2POp(dist.isend, base[plane:2 * plane], lower), dist.P2POp(dist.irecv, base[:plane], lower)]
        if upper is not None:
            ops += [dist.P2POp(dist.isend, base[n:n + plane], upper), dist.P2POp(dist.irecv, base[plane + n:], upper)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    for _ in range(5):
        nccl_exchange()
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for _ in range(200):
        nccl_exchange()
    torch.cuda.synchronize(); t_nccl = (time.perf_counter() - t0) / 200
    ref_lo, ref_hi = base[:plane].clone(), base[plane + n:].clone()
    base[:plane] = 0; base[plane + n:] = 0

    # ---- symmetric-memory mailbox: [slot(2)][direction(2)][plane]
    mb = symm_mem.empty(2 * 2 * plane, dtype=torch.float64, device=dev)
    hdl = symm_mem.rendezvous(mb, dist.group.WORLD)
    peer_lo = hdl.get_buffer(lower, (2 * 2 * plane,), torch.float64) if lower is not None else None
    peer_hi = hdl.get_buffer(upper, (2 * 2 * plane,), torch.float64) if upper is not None else None
    state = {"it": 0}
    def symm_exchange():
        s = state["it"] & 1
        state["it"] += 1
        o = s * 2 * plane
        if peer_lo is not None:   # my bottom plane -> lower neighbour's "from upper" box
            peer_lo[o + plane:o + 2 * plane].copy_(base[plane:2 * plane])
        if peer_hi is not None:   # my top plane -> upper neighbour's "from lower" box
            peer_hi[o:o + plane].copy_(base[n:n + plane])
        hdl.barrier(channel=0)
        if lower is not None:
            base[:plane].copy_(mb[o:o + plane])
        if upper is not None:
            base[plane + n:].copy_(mb[o + plane:o + 2 * plane])
    for _ in range(5):
        symm_exchange()
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for _ in range(200):
        symm_exchange()
    torch.cuda.synchronize(); t_symm = (time.perf_counter() - t0) / 200
    ok = torch.equal(base[:plane], ref_lo) and torch.equal(base[plane + n:], ref_hi)
    if rank == 0:
        print(f"plane {plane * 8 / 1e6:.1f} MB: NCCL {t_nccl * 1e6:.1f} us/exchange, symmetric-memory {t_symm * 1e6:.1f} us/exchange, "
              f"correct={ok}, multicast={hdl.has_multicast_support if hasattr(hdl, 'has_multicast_support') else None}", flush=True)
dist.barrier()
dist.destroy_process_group()
