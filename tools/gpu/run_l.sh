set -x
# (1) launch list of the bench command (gpu__time_duration only; cold-cache, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --config4-streams 200000 > gpurun_out/b_under_ncu.log 2>&1
tail -c 600 gpurun_out/b_under_ncu.log
# (2) ncu --set full of every kernel
ncu --set full --clock-control none --import-source on -k regex:acm_decode_fast2 -s 2 -c 1 -f -o gpurun_out/prof_r02_fast2_c2 python tools/profile_run.py --streams 10000 --runs 3 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:acm_decode_fast2 -s 1 -c 1 -f -o gpurun_out/prof_r02_fast2_c4 python tools/profile_run.py --streams 125000 --runs 2 --workload config4 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:acm_decode_generic -c 1 -f -o gpurun_out/prof_r02_generic python tools/prof_aux.py --streams 20000 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:acmgen_write -c 1 -f -o gpurun_out/prof_r02_gen_write python tools/prof_aux.py --streams 20000 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:acmgen_size -c 1 -f -o gpurun_out/prof_r02_gen_size python tools/prof_aux.py --streams 20000 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:acm_gather_headers -c 1 -f -o gpurun_out/prof_r02_gather python tools/prof_aux.py --streams 20000 2>&1 | tail -1
ls -la gpurun_out/*.ncu-rep | tail -8
