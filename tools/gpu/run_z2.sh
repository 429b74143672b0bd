for d in 0 1; do
echo "== deal $d (F2_PROF build)"
F2_PROF=1 ACM_B200_DEAL=$d ACM_B200_LIB=libacm_b200/_lib/var/prof/libacm_b200.so timeout 300 python tools/profile_run.py --streams 10000 --runs 3 2>&1 | tail -7
done
