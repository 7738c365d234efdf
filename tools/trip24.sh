#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_double.py tests/test_gpu_slabs.py -m gpu -q -x > $O/t24_pytest.log 2>&1; echo "rc=$?" >> $O/t24_pytest.log
B="--steps 400 --warmup 40 --no-cpu-baseline --no-weak --no-e2e"
for a in 0 1; do
  AEROLAB_LBM_AUX2=$a AEROLAB_LBM_DOUBLE=1 timeout 300 python bench.py --workload "configs[4]-case" $B > "$O/t24_c4_aux$a.json" 2> "$O/t24_c4_aux$a.err"
  AEROLAB_LBM_AUX2=$a AEROLAB_LBM_DOUBLE=1 timeout 300 python bench.py --workload "configs[2]" $B > "$O/t24_c2_aux$a.json" 2> "$O/t24_c2_aux$a.err"
done
AEROLAB_LBM_AUX2=0 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-weak --no-e2e > "$O/t24_c3_aux0.json" 2> "$O/t24_c3_aux0.err"
AEROLAB_LBM_AUX2=1 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-weak --no-e2e > "$O/t24_c3_aux1.json" 2> "$O/t24_c3_aux1.err"
for hs in 6 8 12 16 20 24 32; do
  AEROLAB_LBM_S2_HS=$hs AEROLAB_LBM_DOUBLE=1 timeout 300 python bench.py --workload "configs[4]-case" $B > "$O/t24_c4_hs$hs.json" 2> "$O/t24_c4_hs$hs.err"
done
for hs in 16 24; do
  AEROLAB_LBM_S2_HS=$hs AEROLAB_LBM_DOUBLE=1 timeout 300 python bench.py --workload "configs[2]" $B > "$O/t24_c2_hs$hs.json" 2> "$O/t24_c2_hs$hs.err"
done
echo done
