#!/bin/bash
# GPU trip 2 of round 2: cp.async march kernel + verified division by tau
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/t2_pytest.log 2>&1; echo "rc=$?" >> $O/t2_pytest.log
B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e"
$B > $O/t2_c3.json 2> $O/t2_c3.err
for hs in 24 32 48 64 96; do AEROLAB_LBM_S2_HS=$hs $B > $O/t2_c3_hs$hs.json 2>> $O/t2_c3.err; done
AEROLAB_LBM_DIV=ieee AEROLAB_LBM_S2_HS=64 $B > $O/t2_c3_hs64_ieee.json 2>> $O/t2_c3.err
for w in "configs[2]" "configs[4]-case"; do
  n=$(echo $w | tr -d '[]-' )
  python bench.py --workload "$w" --steps 400 --warmup 40 --no-cpu-baseline --no-e2e > $O/t2_${n}_single.json 2> $O/t2_${n}.err
  AEROLAB_LBM_DOUBLE=1 python bench.py --workload "$w" --steps 400 --warmup 40 --no-cpu-baseline --no-e2e > $O/t2_${n}_double.json 2>> $O/t2_${n}.err
done
AEROLAB_LBM_S2_HS=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:march2 -s 3 -c 1 -o $O/r2b_march_c3 \
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > $O/t2_ncu.log 2>&1
echo done
