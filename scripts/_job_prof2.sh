mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --profile --no-cpu-baseline --no-e2e --no-secondary --no-parity > gpurun_out/bench_profile_n2.json 2> gpurun_out/bench_profile_n2.err; echo "bench rc=$?"
grep "^#" gpurun_out/bench_profile_n2.err | head -45
python - <<P
import json
d=json.loads(open("gpurun_out/bench_profile_n2.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["comm"])
P
