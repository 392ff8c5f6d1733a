#!/usr/bin/env python
"""bench.py -- design iterations / s of the compliance hot path (BASELINE.json metric).

A "step" is ONE design iteration of the 3-D cantilever compliance problem on a hex8 grid (default 256x128x128,
12.8 M dof): density filter -> SIMP -> stiffness assembly -> LinSolve (LDAS + CG(tol 1e-8) + geometric multigrid,
warm-started) -> compliance -> adjoint (LDAS, no CG) -> element sensitivities -> SIMP' -> filter^T.
Between steps the design is perturbed, x <- clip(x + 0.2 (rand - 0.5), 0, 1) (tests/bench_assembly.py:80-90 of
the reference), all designs pre-generated from a fixed seed.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size NX NY NZ] [--impl b200|reference]

`value`      device-resident throughput (inputs already in HBM), CUDA-event timed, max over ranks
`e2e`        the same step driven from pinned HOST buffers: x copied host->device and (compliance, dc/dx) copied
             device->host inside the timed region, through the public Module API
`roofline`   dominant kernel (fused damped-Jacobi sweep on the finest level) vs the measured HBM peak
`cpu_baseline` the CPU oracle port of the same step on a bounded sample grid, scaled linearly in dof
--impl reference runs only that CPU arm (all ranks but 0 exit immediately).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

XMIN, RADIUS, TOL = 1e-9, 2.0, 1e-8
METRIC, UNIT = "design_iters_per_sec", "iter/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, nargs=3, default=None, metavar=("NX", "NY", "NZ"))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-size", type=int, nargs=3, default=None, help="sample grid of the CPU arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--problem", default="cantilever", choices=["cantilever", "thermal"],
                    help="cantilever = BASELINE configs[2]/[3] (metric); thermal = configs[4] (scalar conduction heat sink)")
    ap.add_argument("--csr", action="store_true", help="stream the assembled CSR values on every level (no matrix-free level 0)")
    ap.add_argument("--kernel-only", action="store_true", help="only the dominant-kernel loop (for ncu captures)")
    ap.add_argument("--profile", action="store_true", help="per-entry-point CUDA-event breakdown of one step (diagnostic)")
    return ap.parse_args()


def design_sequence(nel, count, seed=1234, keep=None):
    """x_0 = 0.5, then successive bounded random perturbations (seeded).  ``keep`` maps a global design to the part a
    rank stores (its slab); the random stream is always the global one so every rank count sees the same designs."""
    rng = np.random.default_rng(seed)
    keep = (lambda a: a) if keep is None else keep
    x = np.full(nel, 0.5)
    out = [keep(x).copy()]
    for _ in range(count - 1):
        x = np.clip(x + 0.2 * (rng.random(nel) - 0.5), 0.0, 1.0)
        out.append(keep(x).copy())
    return out


# weak scaling: the per-GPU slab stays 256x128x128-equivalent (12.8 M dof) as ranks are added
WEAK_GRIDS = {1: (256, 128, 128), 2: (256, 128, 256), 4: (256, 256, 256), 8: (512, 256, 256)}


# ------------------------------------------------------------------------------------------------ CPU arm (oracle)
def cpu_arm(size, steps, warmup):
    """The reference's algorithm on the host cores (numpy/scipy oracle port; scipy's kernels are single-threaded)."""
    from oracle import Grid
    from oracle.chain import ComplianceProblem

    nx, ny, nz = size
    t0 = time.perf_counter()
    P = ComplianceProblem(Grid(nx, ny, nz), kind="cantilever", radius=RADIUS, xmin=XMIN, tol=TOL)
    setup = time.perf_counter() - t0
    xs = design_sequence(P.grid.nel, warmup + steps)
    times, its = [], []
    cpu0 = wall0 = None
    for i, x in enumerate(xs):
        if i == warmup:
            cpu0, wall0 = time.process_time(), time.perf_counter()
        t0 = time.perf_counter()
        P.response(x)
        P.sensitivity()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            its.append(P.cg.iterations)
    # threads actually busy on average (scipy's sparse kernels are single-threaded; numpy/BLAS helpers may add a few)
    cores = max(1, round((time.process_time() - cpu0) / max(time.perf_counter() - wall0, 1e-9)))
    return dict(sec_per_iter=sum(times) / len(times), setup_s=setup, cg_iterations=its, ndof=P.f.size, compliance=P.c,
                cores=cores)


def pick_cpu_size(full, steps, warmup, explicit):
    if explicit is not None:
        return tuple(explicit)
    # ~2 s / iteration at 64x32x32 and ~9-12 s at 128x64x64 (10 GB resident); keep the whole arm within a few minutes
    try:
        import psutil

        enough_ram = psutil.virtual_memory().available > 24e9
    except Exception:
        enough_ram = False
    if (steps + warmup) * 12 + 60 <= 200 and min(full) >= 64 and enough_ram:
        return (128, 64, 64)
    return (64, 32, 32) if min(full) >= 32 else tuple(full)


def run_reference(args, full):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    size = pick_cpu_size(full, args.steps, args.warmup, args.cpu_size)
    r = cpu_arm(size, args.steps, args.warmup)
    ndof_full = 3 * (full[0] + 1) * (full[1] + 1) * (full[2] + 1)
    scale = r["ndof"] / ndof_full
    value = scale / r["sec_per_iter"]
    sample = (f"oracle port (numpy/scipy) of the same design iteration on a {size[0]}x{size[1]}x{size[2]} grid "
              f"({r['ndof']} dof = {scale:.4f} of the full workload), {args.steps} timed iterations after {args.warmup} "
              f"warm-up, time scaled linearly in dof; CG iterations {r['cg_iterations']}; {os.cpu_count()} host cores present, "
              f"{r['cores']} busy on average (scipy's SpMV / SpGEMM / np.add.at are single-threaded)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * r["sec_per_iter"] / scale, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"3D cantilever compliance {full[0]}x{full[1]}x{full[2]} hex8, CG(1e-8)+GMG, DensityFilter r=2"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock / throttle-reason samples DURING the timed region.  Primary: NVML polled in-process every 5 ms (the timed
    region is a few hundred ms, too short for a freshly spawned ``nvidia-smi -lms`` to report anything); fallback: an
    ``nvidia-smi`` poller that must be started before the warm-up.  Only samples stamped inside [begin(), end()] count."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc, self.nvml, self.h = gpu_index, [], None, None, None
        self.t0 = self.t1 = None
        self._stop = threading.Event()
        self.source = None

    def _nvml_handle(self):
        import pynvml

        pynvml.nvmlInit()
        try:
            import torch

            uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in vis.split(",") if v.strip().isdigit()]
            phys = int(ids[self.gpu]) if self.gpu < len(ids) else self.gpu
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)

    def start(self):
        """Call BEFORE the warm-up steps."""
        try:
            self.nvml, self.h = self._nvml_handle()
            self.max_sm = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.h, self.nvml.NVML_CLOCK_SM))
            self.source = "nvml"
            self.t = threading.Thread(target=self._poll_nvml, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.t = threading.Thread(target=self._read_smi, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        n = self.nvml
        bits = ((n.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"), (n.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (n.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"), (n.nvmlClocksEventReasonSwPowerCap, "sw_power_cap"))
        while not self._stop.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                try:
                    mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((time.perf_counter(), sm, self.max_sm, [nm for b, nm in bits if mask & b]))
            except Exception:
                pass
            time.sleep(0.005)

    def _read_smi(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            if len(r) >= 9 and r[1].replace(".", "").isdigit() and r[2].replace(".", "").isdigit():
                self.rows.append((time.perf_counter(), float(r[1]), float(r[2]),
                                  [nm for nm, v in zip(self.NAMES, r[5:9]) if v.lower().startswith("active")]))

    def begin(self):
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        self._stop.set()
        if self.proc is not None:
            self.proc.terminate()
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi on this host"], "samples": 0}
        self.t.join(timeout=2)
        t0 = self.t0 if self.t0 is not None else -1e300
        t1 = self.t1 if self.t1 is not None else 1e300
        rows = [r for r in self.rows if t0 <= r[0] <= t1]
        scope = "timed region"
        if not rows:  # poller too slow for the region: fall back to everything since start() (warm-up included) and say so
            rows, scope = list(self.rows), "warm-up + timed region"
        reasons = sorted({nm for r in rows for nm in r[3]})
        return {"sm_mhz": statistics.median([r[1] for r in rows]) if rows else None,
                "sm_max_mhz": max(r[2] for r in rows) if rows else None, "reasons": reasons, "samples": len(rows),
                "source": self.source, "scope": scope}


# ------------------------------------------------------------------------------------------------ GPU arm
class GpuChain:
    """The design iteration through the public Module API of pymoto_b200 (device-resident tensors).  With more than one
    rank the grid is split into z-slabs (pymoto_b200/slab.py) and every rank holds its own element layers / node planes."""

    def __init__(self, size, world=1, problem="cantilever"):
        import pymoto_b200 as pmb
        from pymoto_b200 import device as dv

        self.pmb, self.dv = pmb, dv
        nx, ny, nz = size
        self.dom = dom = pmb.VoxelDomain(nx, ny, nz)
        ndof = 3 if problem == "cantilever" else 1
        self.mgs = pmb.solvers.auto_multigrid(dom)
        self.ctx = pmb.slab.init(dom, n_levels=len(self.mgs) + 1, ndof=ndof) if world > 1 else pmb.slab.context(nz)
        k0, k1 = self.ctx.part.planes(0) if world > 1 else (0, nz + 1)
        self.e0, self.e1 = self.ctx.part.elem_layers(0) if world > 1 else (0, nz)
        self.lay = nx * ny
        plane = (nx + 1) * (ny + 1) * ndof
        if problem == "cantilever":
            nodes_face = (np.arange(nz + 1)[:, None] * (ny + 1) + np.arange(ny + 1)[None, :]).ravel() * (nx + 1)  # i = 0
            bc = (nodes_face[:, None] * ndof + np.arange(ndof)[None, :]).ravel()  # global dof numbers
            f = np.zeros((k1 - k0) * plane)  # this rank's node planes
            kl = nz // 2
            if k0 <= kl < k1:
                load_nodes = ((kl - k0) * (ny + 1) + np.arange(ny + 1)) * (nx + 1) + nx  # i = nx, k = nz/2
                f[load_nodes * ndof + 2] = 1.0
        else:  # heat sink: T = 0 on a centred patch of face i = 0, unit heat load on every node with i >= 1
            kk, jj = np.meshgrid(np.arange(nz // 4, (nz + 1) - nz // 4), np.arange(ny // 4, (ny + 1) - ny // 4), indexing="ij")
            bc = ((kk * (ny + 1) + jj) * (nx + 1)).ravel()
            f = np.ones(((k1 - k0), ny + 1, nx + 1))
            f[:, :, 0] = 0.0
            f = f.ravel()
        self.ndof_global = dom.nnodes * ndof
        self.f = dv.to_device(f)
        self.flt = pmb.DensityFilter(dom, radius=RADIUS)
        self.simp = pmb.SIMP(XMIN, 3)
        self.asm = (pmb.AssembleStiffness if problem == "cantilever" else pmb.AssemblePoisson)(dom, bc=np.sort(bc))
        self.cg = pmb.solvers.CG(preconditioner=self.mgs[0], tol=TOL)
        self.ls = pmb.LinSolve(hermitian=True, solver=self.cg)
        self.compl = pmb.Compliance()

    def local(self, x_global):
        return x_global[self.e0 * self.lay:self.e1 * self.lay]

    def step(self, x):
        y = self.flt(x)
        s = self.simp(y)
        K = self.asm(s)
        u = self.ls(K, self.f)
        c = self.compl(u, self.f)
        du, _ = self.compl._sensitivity(1.0)
        dK, _ = self.ls._sensitivity(du)
        ds = self.asm._sensitivity(dK)[0]
        dy = self.simp._sensitivity(ds)
        dx = self.flt._sensitivity(dy)
        return c, dx


def run_b200(args, full):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # keep stdout for the ONE JSON line: anything libraries print (e.g. the NCCL version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as ge

    ge.build()
    from pymoto_b200 import _lib, device as dv

    if args.csr:
        from pymoto_b200.matrix import DeviceCSR as _D

        _D.matrix_free = False
    chain = GpuChain(full, world, args.problem)
    W, K = args.warmup, args.steps
    xs_host = design_sequence(chain.dom.nel, W + K + (1 if args.profile else 0), keep=chain.local)
    nel = xs_host[0].size  # elements held by this rank
    pinned = [torch.from_numpy(x).pin_memory() for x in xs_host]
    xs_dev = [p.to("cuda", non_blocking=True) for p in pinned]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.kernel_only:  # assemble once, then fused Jacobi sweeps on the finest level only
        K0 = chain.asm(chain.simp(chain.flt(xs_dev[0])))
        D0 = K0.diagonal_device()
        va, vb, vc = dv.zeros(K0.shape[0]), dv.empty(K0.shape[0]), dv.to_device(np.random.default_rng(0).random(K0.shape[0]))
        for _ in range(12):
            K0.apply(_lib.JACOBI, va, vb, b=vc, diag=D0, w=0.5)
            va, vb = vb, va
        torch.cuda.synchronize()
        return
    # ---------------- device-resident timing
    sampler = ClockSampler(local)
    sampler.start()
    for i in range(W):
        chain.step(xs_dev[i])
    barrier()
    sampler.begin()
    launches0 = _lib.launch_count
    stats0 = dict(_lib.call_stats)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    cg_its, compl = [], []
    ev[0].record()
    for i in range(K):
        c, dx = chain.step(xs_dev[W + i])
        ev[i + 1].record()
        cg_its.append(chain.cg.iterations)
        compl.append(c)
    barrier()
    sampler.end()
    clocks = sampler.stop()
    launches = _lib.launch_count - launches0
    stats = {k: v - stats0.get(k, 0) for k, v in _lib.call_stats.items() if v - stats0.get(k, 0) > 0}
    total_ms = ev[0].elapsed_time(ev[K])
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(K)]
    # one slab-decomposed job over all ranks.  Weak scaling: the grid grows with the rank count, so the whole-job
    # throughput is quoted in 256x128x128-equivalent design iterations (iterations/s x dof / 12.83 M dof)
    iters_per_sec = K / (total_ms * 1e-3)
    dof_scale = chain.ndof_global / (3 * 257 * 129 * 129) if (world > 1 and args.problem == "cantilever") else 1.0
    value = iters_per_sec * dof_scale
    compl = [float(c) for c in compl]

    if args.profile:
        from pymoto_b200.solvers import GeometricMultigrid as _GMG

        graphs, _GMG.use_cuda_graph = _GMG.use_cuda_graph, False  # time every call individually
        _lib.profile_times = {}
        chain.step(xs_dev[W + K])
        prof, _lib.profile_times = _lib.profile_times, None
        _GMG.use_cuda_graph = graphs
        tot = sum(v[1] for v in prof.values())
        if rank == 0:
            print(f"# per-call breakdown of one step (synchronised calls), total {tot:.2f} ms", file=sys.stderr)
            for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1]):
                print(f"# {v[1]:9.3f} ms {100 * v[1] / tot:5.1f}%  n={v[0]:5d}  avg {1e3 * v[1] / v[0]:9.1f} us  {k}", file=sys.stderr)

    # ---------------- end to end from pinned host buffers (x in, compliance + dc/dx out)
    e2e = None
    if not args.no_e2e:
        chain.ls._u_dev = None  # same cold-start state as the device-resident run's first warm-up step
        out_pinned = torch.empty(nel, dtype=torch.float64).pin_memory()
        c_pinned = torch.empty(1, dtype=torch.float64).pin_memory()
        for i in range(W):
            chain.step(pinned[i].to("cuda", non_blocking=True))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2e_steps = []
        e0.record()
        for i in range(K):
            t_step = time.perf_counter()
            xd = pinned[W + i].to("cuda", non_blocking=True)
            c, dx = chain.step(xd)
            out_pinned.copy_(dx, non_blocking=True)
            c_pinned.copy_(c.reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()  # the user reads the result on the host every step
            e2e_steps.append(1e3 * (time.perf_counter() - t_step))
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        e2e = {"value": dof_scale * K / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * nel * world,
               "d2h_bytes_per_step": (8 * nel + 8) * world, "ms_per_step_list": e2e_steps}

    # ---------------- roofline.  Two kernels carry the step:
    #  (1) the HBM-bound stencil-CSR kernel (tile_kernel, fused Jacobi sweep): timed on the finest ASSEMBLED matrix
    #      (the CSR values of level 0, 8*nnz bytes per sweep) -- this is `roofline`;
    #  (2) with --matrix-free (default) level 0 is applied from the element densities instead (elem_kernel, FP64-pipe
    #      bound, 0.46 GB per application): reported under `matrix_free` with the time the same application would
    #      need at 100 % of the HBM peak if it streamed the assembled values.
    from pymoto_b200.matrix import DeviceCSR

    A = chain.asm._mat
    n, nnz = A.shape[0], A.nnz
    mg0 = chain.mgs[0]
    u, u2, b = mg0._buf["u"], mg0._buf["u2"], mg0._buf["t"]
    D = mg0.smoother.D

    def time_sweeps(reps=20):
        nonlocal u, u2
        for _ in range(3):
            A.apply(_lib.JACOBI, u, u2, b=b, diag=D, w=0.5)
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(reps):
            A.apply(_lib.JACOBI, u, u2, b=b, diag=D, w=0.5)
            u, u2 = u2, u
        k1.record()
        torch.cuda.synchronize()
        return k0.elapsed_time(k1) / reps

    was_mf = DeviceCSR.matrix_free
    DeviceCSR.matrix_free = False
    kern_ms = time_sweeps()
    DeviceCSR.matrix_free = was_mf
    alg_bytes = 8 * nnz + 32 * n  # values once; x, b, diag read and y written once (Jacobi sweep)
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    fine_csr = sum(v for (name, det), v in stats.items() if name == "pmb_spmv" and det[0] == full[0])
    fine_mf = sum(v for (name, det), v in stats.items() if name == "pmb_elem_spmv" and det[0] == full[0])
    csr_calls = sum(v for (name, det), v in stats.items() if name == "pmb_spmv")
    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed `ncu --set full` capture
    # (profiles/ncu_full_tile_kernel_r1b.txt: 8.5216 GB + 0.0906 GB per launch at 256x128x128)
    traffic = 8.5216e9 + 0.0906e9 if (tuple(full) == (256, 128, 128) and world == 1) else None
    step_avg_ms = total_ms / K
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": f"tile_kernel<{A.grid.ndof},JACOBI> on the finest assembled matrix", "kernel_ms": kern_ms,
                "algorithmic_bytes": alg_bytes, "peak_source": peak_src,
                "reference_layout_gbs": (12 * nnz + 4 * (n + 1) + 40 * n) / (kern_ms * 1e-3) / 1e9,
                "fine_level_operator_launches_per_step": (fine_csr + fine_mf) / K,
                "stencil_csr_launches_per_step": csr_calls / K,
                "share_of_step": (fine_csr / K) * kern_ms / step_avg_ms if not was_mf else None}
    # the same kernel on the largest operator it streams inside the default (matrix-free level 0) iteration: level 1
    if len(chain.mgs) > 1 and chain.mgs[0].Ac is not None and chain.mgs[1]._buf is not None:
        A1, mg1 = chain.mgs[0].Ac, chain.mgs[1]
        a1, a2, b1, D1 = mg1._buf["u"], mg1._buf["u2"], mg1._buf["t"], mg1.smoother.D
        for _ in range(3):
            A1.apply(_lib.JACOBI, a1, a2, b=b1, diag=D1, w=0.5)
        torch.cuda.synchronize()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for _ in range(20):
            A1.apply(_lib.JACOBI, a1, a2, b=b1, diag=D1, w=0.5)
            a1, a2 = a2, a1
        q1.record()
        torch.cuda.synchronize()
        l1_ms = q0.elapsed_time(q1) / 20
        l1_bytes = 8 * A1.nnz + 32 * A1.shape[0]
        roofline["level1"] = {"kernel_ms": l1_ms, "algorithmic_bytes": l1_bytes, "achieved": l1_bytes / (l1_ms * 1e-3) / 1e9,
                              "frac": l1_bytes / (l1_ms * 1e-3) / 1e9 / peak,
                              "launches_per_step": sum(v for (nm, det), v in stats.items() if nm == "pmb_spmv" and det[0] == full[0] // 2) / K}
    matrix_free = None
    if was_mf:
        mf_ms = time_sweeps()
        ndof_ = A.grid.ndof
        flops = 2.0 * 8 * (8 * ndof_ * ndof_ + ndof_) * (n / ndof_)  # 8 elements x (8 nodes x ndof^2 + ndof) FMA per node
        matrix_free = {"kernel": f"elem_kernel<{A.grid.ndof},3D,JACOBI> (finest level from element densities)", "kernel_ms": mf_ms,
                       "bound": "fp64 pipe", "fp64_tflops": flops / (mf_ms * 1e-3) / 1e12, "dram_bytes_algorithmic": 40 * n + 8 * nel,
                       "variant": _lib.query("pmb_elem_get_variant", ndof_), "variant_ms_autotune": DeviceCSR.elem_timings_ms.get(ndof_),
                       "launches_per_step": fine_mf / K, "share_of_step": (fine_mf / K) * mf_ms / step_avg_ms,
                       "speedup_vs_streaming_assembled_values": kern_ms / mf_ms,
                       "assembled_layout_time_at_100pct_hbm_peak_ms": alg_bytes / (peak * 1e9) * 1e3,
                       "step_ms_if_level0_streamed_at_100pct_hbm_peak": step_avg_ms + (fine_mf / K) * (alg_bytes / (peak * 1e9) * 1e3 - mf_ms)}

    # ---------------- the same K steps with every level streamed from its assembled CSR values (north-star layout)
    csr_streamed = None
    if was_mf and world == 1 and not args.no_e2e:
        DeviceCSR.matrix_free = False
        chain.ls._u_dev = None
        for i in range(W):
            chain.step(xs_dev[i])
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for i in range(K):
            chain.step(xs_dev[W + i])
        c1.record()
        torch.cuda.synchronize()
        DeviceCSR.matrix_free = True
        cms = c0.elapsed_time(c1) / K
        csr_streamed = {"value": 1e3 / cms, "unit": UNIT, "ms_per_step": cms,
                        "fine_kernel_share_of_step": (fine_mf / K) * kern_ms / cms,
                        "note": "same workload with DeviceCSR.matrix_free = False (bench.py --csr)"}

    # ---------------- CPU baseline (bounded sample), rank 0 only
    cpu = None
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        size = tuple(args.cpu_size) if args.cpu_size else ((64, 32, 32) if min(full) >= 32 else tuple(full))
        r = cpu_arm(size, 3, 1)
        scale = r["ndof"] / n
        cpu = {"value": scale / r["sec_per_iter"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": f"oracle port on {size[0]}x{size[1]}x{size[2]} ({r['ndof']} dof), 3 timed iterations after 1 warm-up, "
                         f"{r['sec_per_iter']:.3f} s/iteration, scaled linearly in dof ({scale:.5f}); CG its {r['cg_iterations']}; "
                         f"scipy SpMV/SpGEMM and np.add.at are single-threaded ({os.cpu_count()} host cores present)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"3D {'cantilever compliance' if args.problem == 'cantilever' else 'heat-sink (scalar conduction) compliance'} "
                                   f"{full[0]}x{full[1]}x{full[2]} hex8 ({chain.ndof_global} dof, "
                                   f"{n} dof / nnz {nnz} per GPU), "
                                   f"SIMP p=3 xmin=1e-9, DensityFilter r=2, LDAS+CG(tol 1e-8)+GMG({len(chain.mgs)} levels, "
                                   "5+5 Jacobi w=0.5, V-cycle replayed as one CUDA graph on 1 GPU), warm start, seeded design perturbations; "
                                   "finest-level operator " +
                                   ("matrix-free (element-wise)" if not args.csr else "streamed from the assembled CSR values"),
                       "l2": "inputs larger than L2 (matrix values 8*nnz bytes per level-0 sweep)",
                       "parallelism": "1 GPU" if world == 1 else
                       f"{world} z-slabs ({chain.ctx.part.n_dist} split multigrid levels, coarser levels replicated), NCCL halo "
                       "exchange + dot-product all-reduce; value = iterations/s x dof / 12.83M (weak scaling)",
                       "iters_per_sec_this_grid": iters_per_sec},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "matrix_free": matrix_free,
            "csr_streamed": csr_streamed, "cpu_baseline": cpu,
            "cg_iterations": cg_its, "ms_per_step_list": step_ms, "compliance": compl,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    full = tuple(args.size) if args.size else WEAK_GRIDS.get(world, (256, 128, 128 * world))
    if args.impl == "reference":
        run_reference(args, full)
    else:
        run_b200(args, full)


if __name__ == "__main__":
    main()
