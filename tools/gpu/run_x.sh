set -x
ncu --set full --clock-control none --import-source on -k regex:acm_unpack -c 1 -o gpurun_out/prof_r02_split_unpack_c4 -f python tools/profile_run.py --streams 125000 --runs 1 --workload config4 --kernel 2 2>&1 | tail -1
