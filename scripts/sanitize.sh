#!/bin/bash
# compute-sanitizer passes over the small-grid GPU tests (run on the GPU box; results summarised in profiles/sanitizer_*.txt).
#   gpurun --timeout 1500 -- 'bash scripts/sanitize.sh > gpurun_out/sanitize.log 2>&1'
# Round 2: the parity-block matrix-free layouts (TMA bulk-copy ring, shared-memory carries / exchange), the symmetric
# half-stencil kernels, the transfer kernels, the direct Galerkin build, the MMA passes and the C PCG driver.
set -u
K='parity_block or symmetric_half or variants_bit_identical or galerkin_direct or transfer_and or mma_device_update or c_pcg or restriction_constants'
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $tool -k \"$K\""
  compute-sanitizer --tool "$tool" --error-exitcode 1 python -m pytest tests -m gpu -q -x -k "$K" 2>&1 | tail -12
done
