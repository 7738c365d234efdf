#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
python examples/polar_sweep.py --steps 20000 --alpha-min 2 --alpha-max 2 --out $O/tmp_polar.csv > /dev/null 2>&1   # page-in
for r in 1 2; do
python examples/polar_sweep.py --steps 20000 --alpha-min 2 --alpha-max 5 --out $O/r2h_polar4_naca0012_2048x1024_20000steps.csv 2>> $O/t32.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('4 cases', d['seconds'], d['aggregate_glups'])" >> $O/t32.log
done
python examples/polar_sweep.py --steps 20000 --alpha-min 2 --alpha-max 2 --out $O/tmp_polar.csv 2>> $O/t32.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1 case', d['seconds'], d['aggregate_glups'])" >> $O/t32.log
python examples/polar_sweep.py --steps 20000 --alpha-min 2 --alpha-max 9 --out $O/tmp_polar.csv 2>> $O/t32.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('8 cases', d['seconds'], d['aggregate_glups'])" >> $O/t32.log
echo done
