set -x
for k in walk unpack lift; do
ncu --set full --clock-control none --import-source on -k regex:acm_$k -c 1 -o gpurun_out/prof_r02_split_${k}_c4 -f python tools/profile_run.py --streams 125000 --runs 1 --workload config4 --kernel 2 2>&1 | tail -2
done
ncu --set full --clock-control none --import-source on -k regex:acm_walk -c 1 -o gpurun_out/prof_r02_split_walk_c2 -f python tools/profile_run.py --streams 10000 --runs 1 --kernel 2 2>&1 | tail -2
ls -la gpurun_out/*.ncu-rep
