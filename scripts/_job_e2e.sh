mkdir -p gpurun_out
timeout 500 python bench.py --steps 20 --no-cpu-baseline --no-secondary --no-parity > gpurun_out/bench_e2e20.json 2> gpurun_out/bench_e2e20.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads(open("gpurun_out/bench_e2e20.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], [round(v,1) for v in d["e2e"]["ms_per_step_list"]], d["e2e"].get("device_allocs_in_timed_region"))
P
