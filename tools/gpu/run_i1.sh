python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-900
