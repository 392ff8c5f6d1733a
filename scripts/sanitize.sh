#!/bin/bash
# compute-sanitizer passes over the small-grid GPU tests (run on the GPU box; results summarised in profiles/sanitizer_*.txt).
# Round 1 covered the kernels that existed at the time (profiles/sanitizer_r1.txt); the z-marching / tensor-core matrix-free
# layouts, the MMA passes and the C PCG driver were added later and still need a pass:
#   gpurun --timeout 900 -- 'bash scripts/sanitize.sh > gpurun_out/sanitize.log 2>&1'
set -u
K='variants or mma_device_update or c_pcg or write_to_vti or assembly_vs_oracle_ragged or filter_vs_oracle_ragged or restriction_constants'
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $tool"
  compute-sanitizer --tool "$tool" --error-exitcode 1 python -m pytest tests -m gpu -q -x -k "$K" 2>&1 | tail -15
done
