B="python bench.py --workload config3 --steps 1 --warmup 0 --no-cpu --no-config4 --no-e2e --no-streaming"
for k in acm_scan acm_unpack_any acm_lift_tile; do
ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/prof_r02_g3_${k#acm_} -f $B 2>&1 | tail -1
done
ncu --set full --clock-control none --import-source on -k regex:acm_walk1 -s 2 -c 1 -o gpurun_out/prof_r02_walk1 -f python tools/dbg_stream.py 2>&1 | tail -1
