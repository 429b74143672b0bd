# grouped general path: sanitizers, ncu capture of the new scan kernel, launch list and bench record of config 3
for t in memcheck racecheck; do
timeout 600 compute-sanitizer --tool $t python tools/sanitize_groups.py 2>&1 | grep -v '^=========     Host Frame\|^=========         in ' | tail -3
done
B="python bench.py --workload config3 --steps 1 --warmup 0 --no-cpu --no-config4 --no-e2e --no-streaming"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:acm_scan -c 1 -o gpurun_out/prof_r02_g3_scan_v2 -f $B 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_config3.csv python bench.py --workload config3 --steps 2 --warmup 3 --no-cpu --no-config4 --no-e2e --no-streaming > /dev/null 2>&1
tail -5 gpurun_out/launches_r02_config3.csv | cut -c1-200
timeout 900 python bench.py --workload config3 --no-config4 > gpurun_out/bench_r02d_config3.json 2> gpurun_out/bench_r02d_config3.err
tail -c 900 gpurun_out/bench_r02d_config3.json
