timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -30 > gpurun_out/r2s_tests.log
tail -4 gpurun_out/r2s_tests.log
timeout 200 python scripts/time_elem.py --out gpurun_out/time_elem_r2s.json > gpurun_out/time_elem_r2s.log 2>&1; echo $(grep -E "\"ms\"" gpurun_out/time_elem_r2s.json | tr -d ' \n')
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2s.json 2> gpurun_out/bench_r2s.err; tail -c 300 gpurun_out/bench_r2s.err; head -c 300 gpurun_out/bench_r2s.json
