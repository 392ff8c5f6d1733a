# round-2 late rows: the whole GPU suite (GCMMA, several element matrices, add_constant included) and the default bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2c.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu_r2c.log
timeout 500 python bench.py > gpurun_out/bench_r2c_n1.json 2> gpurun_out/bench_r2c_n1.err; echo "bench rc=$?"
python - <<P
import json
try:
    d=json.loads(open("gpurun_out/bench_r2c_n1.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["parity"], d["roofline"]["frac"], d["cpu_baseline"]["value"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_r2c_n1.err").read()[-2500:])
P
