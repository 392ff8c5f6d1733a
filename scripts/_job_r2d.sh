mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2d.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_r2d.log
timeout 500 python bench.py > gpurun_out/bench_r2d_n1.json 2> gpurun_out/bench_r2d_n1.err; echo "bench rc=$?"
python - <<P
import json
try:
    d=json.loads(open("gpurun_out/bench_r2d_n1.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["parity"]["pass"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["roofline"]["layout_ms_autotune"], d["secondary"]["configs4_thermal"]["ms_per_step"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_r2d_n1.err").read()[-2500:])
P
