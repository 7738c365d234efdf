#!/bin/bash
# GPU trip 1 of round 2: parity of the new kernel, A/B against the ring kernel, ncu of march2_kernel
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/t1_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_large.py > $O/t1_pytest_small.log 2>&1; echo "small rc=$?" >> $O/t1_pytest_small.log
timeout 900 python -m pytest tests/test_gpu_large.py -q > $O/t1_pytest_large.log 2>&1; echo "large rc=$?" >> $O/t1_pytest_large.log
B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e"
$B > $O/t1_c3_march.json 2> $O/t1_c3_march.err
AEROLAB_LBM_S2_KERNEL=ring $B > $O/t1_c3_ring.json 2> $O/t1_c3_ring.err
for hs in 64 128 256; do AEROLAB_LBM_S2_HS=$hs $B > $O/t1_c3_march_hs$hs.json 2>> $O/t1_c3_march.err; done
for w in "configs[2]" "configs[4]-case"; do
  n=$(echo $w | tr -d '[]-' )
  python bench.py --workload "$w" --steps 400 --warmup 40 --no-cpu-baseline --no-e2e > $O/t1_${n}_single.json 2> $O/t1_${n}.err
  AEROLAB_LBM_DOUBLE=1 python bench.py --workload "$w" --steps 400 --warmup 40 --no-cpu-baseline --no-e2e > $O/t1_${n}_double.json 2>> $O/t1_${n}.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march2 -s 3 -c 1 -o $O/r2a_march_c3 \
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > $O/t1_ncu.log 2>&1
echo done
