for v in base s18 s20 s24 w24 w32 w36 base; do
echo "== $v"
if [ $v = base ]; then L=libacm_b200/_lib/libacm_b200.so; else L=libacm_b200/_lib/var/$v/libacm_b200.so; fi
ACM_B200_LIB=$L timeout 300 python tools/profile_run.py --streams 10000 --runs 4 2>&1 | tail -1
ACM_B200_LIB=$L timeout 300 python tools/profile_run.py --streams 125000 --runs 3 --workload config4 2>&1 | tail -1
done
