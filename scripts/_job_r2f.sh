mkdir -p gpurun_out
timeout 500 python bench.py > gpurun_out/bench_r2f_n1.json 2> gpurun_out/bench_r2f_n1.err; echo "bench rc=$?"
python - <<P
import json
try:
    d=json.loads(open("gpurun_out/bench_r2f_n1.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step_list"], d["parity"]["pass"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["secondary"]["configs4_thermal"]["ms_per_step"], d["roofline"]["csr_level1"]["frac"], d["roofline"]["iteration"]["frac"], d["cpu_baseline"]["value"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_r2f_n1.err").read()[-2500:])
P
timeout 200 python -m pytest tests/test_gpu_parity.py -q -x -k "matrix_free or config4 or thermal" 2>&1 | tail -2
