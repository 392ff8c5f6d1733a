# one multi-GPU box: dist_check with the one-launch peer exchange kernels, then bench A/B (fused / unfused / deeper split)
N=$1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
mkdir -p gpurun_out
timeout 600 $T 29511 tests/dist_check.py > gpurun_out/dist_check_n${N}_fused.log 2>&1; echo "dist_check rc=$?"
grep "dist_check\|Error\|error\|assert" gpurun_out/dist_check_n${N}_fused.log | tail -14
run() { # name, env...
  name=$1; shift
  env "$@" timeout 400 $T 29512 bench.py --gpus $N --no-secondary --no-e2e $EXTRA > gpurun_out/bench_n${N}_$name.json 2> gpurun_out/bench_n${N}_$name.err; echo "bench $name rc=$?"
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/bench_n${N}_$name.json").read().strip().splitlines()[-1])
    print("$name", d["ms_per_step"], d["comm"], d["config"]["parallelism"][:60], d["cg_iterations"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/bench_n${N}_$name.err").read()[-1500:])
P
}
EXTRA=--profile run fused PMB_X=1
grep "^#" gpurun_out/bench_n${N}_fused.err | head -30
EXTRA= run unfused PMB_PEER_FUSED=0
EXTRA= run split3 PMB_SLAB_MIN_DOFS=300000
EXTRA= run nccl_allreduce PMB_PEER_ALLREDUCE=0
