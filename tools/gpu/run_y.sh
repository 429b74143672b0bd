# scan-warp pairing (ACM_B200_DEAL): parity of the fused kernel, config 2 and config-4 shape with both deals; config 3 with the refined walk proxy
timeout 900 python -m pytest tests/test_gpu_batch.py -x -q -m gpu -k "fallout or config2_full or more_streams or truncations or fast_shape or unaligned or healthy" 2>&1 | tail -3
for d in 0 1 0 1; do
echo "== deal $d"
ACM_B200_DEAL=$d timeout 300 python tools/profile_run.py --streams 10000 --runs 4 2>&1 | tail -3
done
for d in 0 1; do
echo "== config-4 shape, deal $d"
ACM_B200_DEAL=$d timeout 300 python tools/profile_run.py --streams 125000 --runs 3 --workload config4 2>&1 | tail -2
done
echo "== config 3"
timeout 300 python tools/profile_run.py --streams 10000 --runs 3 --workload config3 2>&1 | tail -2
