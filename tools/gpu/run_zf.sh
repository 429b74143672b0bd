for v in nopair run2 run3 run6 run8 nopair; do
echo "== $v"
ACM_B200_LIB=libacm_b200/_lib/var/$v/libacm_b200.so timeout 300 python tools/profile_run.py --streams 10000 --runs 3 --workload config3 2>&1 | tail -2 | head -1
ACM_B200_LIB=libacm_b200/_lib/var/$v/libacm_b200.so timeout 600 python tools/prof_aux.py --streams 20000 --kernel 1 2>&1 | tail -1
done
