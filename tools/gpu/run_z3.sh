for ng in 8 5; do
for pw in 1.6 2.0 2.6; do
echo "== groups $ng pow $pw"
ACM_B200_GEN_GROUPS=$ng ACM_B200_GEN_GROUP_POW=$pw timeout 300 python tools/profile_run.py --streams 10000 --runs 3 --workload config3 2>&1 | tail -2 | head -1
done
done
