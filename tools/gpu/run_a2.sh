for v in plainld; do
ACM_B200_LIB=libacm_b200/_lib/var/$v/libacm_b200.so ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__t_sector_hit_rate.pct --clock-control none -k regex:"acm_walk1" -c 2 python tools/stream_time.py 2>&1 | grep -E "gpu__time|inst_exec|hit_rate" | head -6
done
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__t_sector_hit_rate.pct --clock-control none -k regex:"acm_walk1" -c 2 python tools/stream_time.py 2>&1 | grep -E "gpu__time|inst_exec|hit_rate" | head -6
