timeout 1500 python bench.py > gpurun_out/bench_r02b_n1.json 2> gpurun_out/bench_r02b_n1.err
tail -c 3000 gpurun_out/bench_r02b_n1.json
timeout 900 python bench.py --workload config3 --no-config4 > gpurun_out/bench_r02b_config3.json 2> gpurun_out/bench_r02b_config3.err
tail -c 2500 gpurun_out/bench_r02b_config3.json
