#!/bin/bash
# GPU trip 13 (1 GPU): DIAG arg-max search skips identical candidates; ensemble single vs double steps
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > $O/t13_pytest.log 2>&1; echo "rc=$?" >> $O/t13_pytest.log
python bench.py --steps 200 --warmup 20 > $O/t13_c3.json 2> $O/t13_c3.err
python bench.py --workload "configs[2]" --steps 2000 --warmup 200 --no-cpu-baseline > $O/t13_c2.json 2> $O/t13_c2.err
python examples/polar_sweep.py --steps 20000 --alpha-min 0 --alpha-max 3 --out $O/t13_polar4.csv > $O/t13_polar4_single.json 2> $O/t13_polar4.err
AEROLAB_LBM_DOUBLE=1 python examples/polar_sweep.py --steps 20000 --alpha-min 0 --alpha-max 3 --out $O/t13_polar4d.csv > $O/t13_polar4_double.json 2>> $O/t13_polar4.err
echo done
