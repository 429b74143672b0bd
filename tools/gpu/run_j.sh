set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02_a.json 2> gpurun_out/bench_r02_a.err; tail -c 3000 gpurun_out/bench_r02_a.json; tail -5 gpurun_out/bench_r02_a.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_a_ref.json 2>> gpurun_out/bench_r02_a.err; tail -c 1500 gpurun_out/bench_r02_a_ref.json
