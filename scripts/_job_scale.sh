# final-code multi-GPU record: dist_check + the full default bench line at N ranks
N=$1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
mkdir -p gpurun_out
timeout 600 $T 29511 tests/dist_check.py > gpurun_out/dist_check_n${N}_final.log 2>&1; echo "dist_check rc=$?"
grep "dist_check\] \(OK\|peer\|[0-9]\)\|Error\|error\|assert" gpurun_out/dist_check_n${N}_final.log | tail -8
timeout 600 $T 29512 bench.py --gpus $N > gpurun_out/bench_n${N}_final.json 2> gpurun_out/bench_n${N}_final.err; echo "bench rc=$?"
python - <<P
import json
try:
    d=json.loads(open("gpurun_out/bench_n${N}_final.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["comm"], d["cg_iterations"], d["e2e"]["value"], d["parity"]["pass"])
    print(json.dumps(d.get("secondary"))[:1500])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_n${N}_final.err").read()[-2500:])
P
