timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -30 > gpurun_out/r2z_tests.log
tail -4 gpurun_out/r2z_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2z.json 2> gpurun_out/bench_r2z.err; tail -c 300 gpurun_out/bench_r2z.err; head -c 400 gpurun_out/bench_r2z.json
