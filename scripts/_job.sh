timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "parity_block or variants_bit_identical" 2>&1 | tail -15 > gpurun_out/r2v_tests.log
tail -3 gpurun_out/r2v_tests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"elem_kernel_par" -s 3 -c 1 -o gpurun_out/prof_par3_r2b python scripts/time_elem.py --variants 8 --reps 2 --cases 3:256x128x128 > gpurun_out/ncu_par3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"elem_kernel_par" -s 3 -c 1 -o gpurun_out/prof_par3_r2c python scripts/time_elem.py --variants 10 --reps 2 --cases 3:256x128x128 > gpurun_out/ncu_par3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"elem_kernel_par" -s 3 -c 1 -o gpurun_out/prof_par1_r2b python scripts/time_elem.py --variants 11 --reps 2 --cases 1:256x256x256 > gpurun_out/ncu_par1.log 2>&1
ls -la gpurun_out/prof_par*
