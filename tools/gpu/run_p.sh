set -x
timeout 900 python -m pytest tests/test_gpu_batch.py -x -q -m gpu 2>&1 | tail -3
M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct"
for v in base nokeep; do
echo "== $v"
L=libacm_b200/_lib/var/$v/libacm_b200.so; [ $v = base ] && L=libacm_b200/_lib/libacm_b200.so
ACM_B200_LIB=$L timeout 300 python tools/profile_run.py --streams 10000 --runs 4 2>&1 | tail -2
ACM_B200_LIB=$L timeout 600 python tools/profile_run.py --streams 125000 --runs 3 --workload config4 2>&1 | tail -1
ACM_B200_LIB=$L ncu $M --clock-control none -k regex:acm_decode_fast2 -s 2 -c 1 python tools/profile_run.py --streams 10000 --runs 3 2>&1 | grep -E "dram__|inst_exec|duration|hit_rate"
ACM_B200_LIB=$L ncu $M --clock-control none -k regex:acm_decode_fast2 -s 1 -c 1 python tools/profile_run.py --streams 125000 --runs 2 --workload config4 2>&1 | grep -E "dram__|inst_exec|duration|hit_rate"
done
