M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct"
for v in p16 nokeep; do
echo "== $v"
ACM_B200_LIB=libacm_b200/_lib/var/$v/libacm_b200.so timeout 600 ncu $M --clock-control none -k regex:acm_decode_fast2 -s 2 -c 1 python tools/profile_run.py --streams 10000 --runs 3 2>&1 | grep -E "dram__|duration|hit_rate"
done
echo "== base, first launch of a fresh process (-s 0)"
timeout 600 ncu $M --clock-control none -k regex:acm_decode_fast2 -s 0 -c 1 python tools/profile_run.py --streams 10000 --runs 1 2>&1 | grep -E "dram__|duration|hit_rate"
