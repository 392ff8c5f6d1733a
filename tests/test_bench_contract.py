"""bench.py contract on CPU: the reference arm (the unmodified reference timed on the host cores; the oracle port only when
the reference cannot be imported) prints exactly one JSON line with the keys the driver reads, and non-zero ranks of a
multi-rank launch do no work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "16", "8", "8", "--steps", "1", "--warmup", "1"]
    return subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300, cwd=ROOT)


def test_reference_arm_json_line():
    res = _run()
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "design_iters_per_sec" and d["unit"] == "iter/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import refload

    have_ref = refload.reference_root() is not None
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert ("UNMODIFIED pyMOTO" if have_ref else "oracle port") in d["cpu_baseline"]["sample"]
    assert d["config"]["same_config"] is False and d["config"]["extrapolation_factor"] >= 1.0
    assert d["e2e"] == {"value": d["value"], "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    res = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_weak_scaling_grids_and_design_sequence():
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench

    assert bench.WEAK_GRIDS[1] == (256, 128, 128) and bench.WEAK_GRIDS[8] == (512, 256, 256)
    for n, (nx, ny, nz) in bench.WEAK_GRIDS.items():  # per-GPU share of the dofs stays ~ the 1-GPU problem
        per = 3 * (nx + 1) * (ny + 1) * (nz + 1) / n
        assert abs(per / (3 * 257 * 129 * 129) - 1) < 0.02
        assert nz % n == 0 and (nz // n) % 2 == 0
    full = bench.design_sequence(1000, 4)
    part = bench.design_sequence(1000, 4, keep=lambda a: a[200:300])
    assert all(np.array_equal(f[200:300], p) for f, p in zip(full, part))  # every rank sees the same global designs
    assert np.all(full[0] == 0.5) and all(0.0 <= f.min() and f.max() <= 1.0 for f in full)
