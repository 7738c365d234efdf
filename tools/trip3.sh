#!/bin/bash
# GPU trip 3 of round 2: slab control surface tests, CTA generations / aux priority, kernel variants
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/t3_pytest.log 2>&1; echo "rc=$?" >> $O/t3_pytest.log
B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e"
$B > $O/t3_c3.json 2> $O/t3_c3.err
AEROLAB_LBM_AUX_PRIO=0 $B > $O/t3_c3_prio0.json 2>> $O/t3_c3.err
AEROLAB_LBM_S2_GENERATIONS=1 $B > $O/t3_c3_gen1.json 2>> $O/t3_c3.err
AEROLAB_LBM_S2_GENERATIONS=12 $B > $O/t3_c3_gen12.json 2>> $O/t3_c3.err
for v in gb2 l2h1 l2h2 w18 w18gb2 gb2l2h1; do
  AEROLAB_LBM_LIB=$PWD/variants/v_$v.so $B > $O/t3_c3_$v.json 2>> $O/t3_c3.err
done
AEROLAB_LBM_S2_HS=64 AEROLAB_LBM_LIB=$PWD/variants/v_l2h1.so $B > $O/t3_c3_l2h1_hs64.json 2>> $O/t3_c3.err
for w in "configs[2]" "configs[4]-case"; do
  n=$(echo $w | tr -d '[]-' )
  AEROLAB_LBM_DOUBLE=1 python bench.py --workload "$w" --steps 400 --warmup 40 --no-cpu-baseline --no-e2e > $O/t3_${n}_double.json 2> $O/t3_${n}.err
  AEROLAB_LBM_DOUBLE=1 AEROLAB_LBM_AUX_PRIO=0 python bench.py --workload "$w" --steps 400 --warmup 40 --no-cpu-baseline --no-e2e > $O/t3_${n}_double_prio0.json 2>> $O/t3_${n}.err
done
# the default command, complete line (e2e, e2e_fields, cpu baseline)
python bench.py > $O/t3_default_line.json 2> $O/t3_default_line.err
python bench.py --impl reference --steps 20 --warmup 3 > $O/t3_reference_line.json 2> $O/t3_reference_line.err
echo done
