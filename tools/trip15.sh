#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march2 -s 0 -c 1 -o $O/r2e_march_diag_c3 \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > $O/t15_ncu.log 2>&1
echo done
