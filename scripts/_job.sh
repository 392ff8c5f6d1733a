timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "matrix_free or design_iteration or golden" 2>&1 | tail -15 > gpurun_out/r2l_tests.log
tail -5 gpurun_out/r2l_tests.log
for sk in 0 1 15; do PMB_YM_SKEW=$sk timeout 200 python scripts/time_elem.py --cases 3:256x128x128 --variants 0,6,7 --out gpurun_out/time_elem_r2l_$sk.json > gpurun_out/time_elem_r2l_$sk.log 2>&1; echo "dbg $sk" $(grep -E "\"ms\"|maxdiff" gpurun_out/time_elem_r2l_$sk.json | tr -d ' \n'); done
tail -3 gpurun_out/time_elem_r2l_0.log
