B="python bench.py --workload config3 --steps 1 --warmup 0 --no-cpu --no-config4 --no-e2e --no-streaming"
for k in acm_scan acm_unpack_any; do
ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/prof_r02_g3_${k#acm_} -f $B 2>&1 | tail -1
done
timeout 900 python bench.py --workload config3 --no-config4 > gpurun_out/bench_r02b_config3.json 2> gpurun_out/bench_r02b_config3.err
tail -c 600 gpurun_out/bench_r02b_config3.json
timeout 600 python tools/prof_aux.py --streams 20000 --kernel 1 2>&1 | tail -1
timeout 600 python tools/prof_aux.py --streams 125000 --kernel 1 2>&1 | tail -1
