timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary --no-parity > gpurun_out/b_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"elem_kernel<" -s 20 -c 1 -o gpurun_out/prof_elem_r2 python bench.py --kernel-only > gpurun_out/ncu_elem_r2.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"assemble_slot|galerkin_direct" -c 2 -o gpurun_out/prof_asm_gal_r2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary --no-parity > gpurun_out/ncu_asm_r2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3; wc -l gpurun_out/launches_r2.csv
