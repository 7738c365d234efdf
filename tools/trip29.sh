#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
P="python examples/polar_sweep.py --steps 20000 --alpha-min 2 --alpha-max 5 --out $O/tmp_polar.csv"
for cfg in "AEROLAB_LBM_AUX2=1 AEROLAB_LBM_MARCH_EDGES=1" "AEROLAB_LBM_AUX2=0 AEROLAB_LBM_MARCH_EDGES=1" "AEROLAB_LBM_AUX2=1 AEROLAB_LBM_MARCH_EDGES=0" "AEROLAB_LBM_AUX2=0 AEROLAB_LBM_MARCH_EDGES=0" "AEROLAB_LBM_DOUBLE=0" "AEROLAB_LBM_AUX_PRIO=0" "AEROLAB_LBM_S2_GENERATIONS=1"; do
  echo "== $cfg" >> $O/t29.log
  env $cfg $P 2>> $O/t29.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['seconds'], d['aggregate_glups'])" >> $O/t29.log
done
# one case alone, and two
echo "== one case" >> $O/t29.log
python examples/polar_sweep.py --steps 20000 --alpha-min 2 --alpha-max 2 --out $O/tmp_polar.csv 2>> $O/t29.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['seconds'], d['aggregate_glups'])" >> $O/t29.log
echo done
