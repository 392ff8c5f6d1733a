mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "filter" > gpurun_out/pytest_filter.log 2>&1; echo "pytest filter rc=$?"; tail -3 gpurun_out/pytest_filter.log
timeout 400 python bench.py --profile --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/bench_profile.json 2> gpurun_out/bench_profile.err; echo "bench rc=$?"
tail -80 gpurun_out/bench_profile.err
python - <<P
import json
d=json.loads(open("gpurun_out/bench_profile.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"])
P
