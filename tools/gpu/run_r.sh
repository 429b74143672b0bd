set -x
for v in base scan27 scan30; do
echo "== $v"
L=libacm_b200/_lib/var/$v/libacm_b200.so; [ $v = base ] && L=libacm_b200/_lib/libacm_b200.so
ACM_B200_LIB=$L timeout 300 python tools/profile_run.py --streams 10000 --runs 4 2>&1 | tail -2
ACM_B200_LIB=$L timeout 600 python tools/profile_run.py --streams 125000 --runs 3 --workload config4 2>&1 | tail -1
done
