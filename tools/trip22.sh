#!/bin/bash
# edges in march2_kernel: parity tests, then A/B bench lines
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_double.py tests/test_gpu_large.py -m gpu -q -x > $O/t22_pytest.log 2>&1; echo "rc=$?" >> $O/t22_pytest.log
for e in 0 1; do
  for w in "configs[4]-case" "configs[2]"; do
    AEROLAB_LBM_MARCH_EDGES=$e AEROLAB_LBM_DOUBLE=1 timeout 300 python bench.py --workload "$w" --steps 400 --warmup 40 --no-cpu-baseline --no-weak > "$O/t22_${w}_e$e.json" 2> "$O/t22_${w}_e$e.err"
  done
  AEROLAB_LBM_MARCH_EDGES=$e timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-weak > "$O/t22_c3_e$e.json" 2> "$O/t22_c3_e$e.err"
done
AEROLAB_LBM_DOUBLE=0 timeout 300 python bench.py --workload "configs[4]-case" --steps 400 --warmup 40 --no-cpu-baseline --no-weak > "$O/t22_c4_single.json" 2> "$O/t22_c4_single.err"
echo done
