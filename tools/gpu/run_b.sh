set -x
ncu --set full --clock-control none --import-source on -k regex:acm_decode_fast2 -s 1 -c 1 -f -o gpurun_out/prof_r02_a_c4 python tools/profile_run.py --streams 40000 --runs 2 --workload config4 2>&1 | tail -4
ncu --set full --clock-control none --import-source on -k regex:acm_decode_fast2 -s 1 -c 1 -f -o gpurun_out/prof_r02_a_c2 python tools/profile_run.py --streams 10000 --runs 2 2>&1 | tail -4
