set -x
ncu --set full --clock-control none --import-source on -k regex:acm_walk -c 1 -o gpurun_out/prof_r02_split_walk_c2 -f python tools/profile_run.py --streams 10000 --runs 1 --kernel 2 2>&1 | tail -1
