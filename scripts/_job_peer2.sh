N=$1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
mkdir -p gpurun_out
timeout 300 $T 29530 scripts/probe_peer_xchg.py 2>&1 | grep -v "Warning\|^$\|\*\*\*\|OMP_NUM" | tee gpurun_out/probe_peer_xchg_n$N.txt
timeout 600 $T 29511 tests/dist_check.py > gpurun_out/dist_check_n${N}_fused.log 2>&1; echo "dist_check rc=$?"
grep "dist_check\|Error\|error\|assert" gpurun_out/dist_check_n${N}_fused.log | tail -14
run() { # name, env...
  name=$1; shift
  env "$@" timeout 400 $T 29512 bench.py --gpus $N --no-secondary --no-e2e > gpurun_out/bench_n${N}_$name.json 2> gpurun_out/bench_n${N}_$name.err; echo "bench $name rc=$?"
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/bench_n${N}_$name.json").read().strip().splitlines()[-1])
    print("$name", d["ms_per_step"], d["comm"], d["config"]["parallelism"][:60], d["cg_iterations"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/bench_n${N}_$name.err").read()[-1500:])
P
}
run fused PMB_X=1
run split3 PMB_SLAB_MIN_DOFS=200000
run unfused_split3 PMB_PEER_FUSED=0 PMB_SLAB_MIN_DOFS=200000
