#!/bin/bash
# GPU trip 12 (1 GPU): DIAG riding inside the collision; ensemble with double steps; full suite
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > $O/t12_pytest.log 2>&1; echo "rc=$?" >> $O/t12_pytest.log
python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/t12_c3.json 2> $O/t12_c3.err
python examples/polar_sweep.py --steps 4000 --alpha-min 0 --alpha-max 3 --out $O/t12_polar4.csv > $O/t12_polar4_single.json 2> $O/t12_polar4.err
AEROLAB_LBM_DOUBLE=1 python examples/polar_sweep.py --steps 4000 --alpha-min 0 --alpha-max 3 --out $O/t12_polar4d.csv > $O/t12_polar4_double.json 2>> $O/t12_polar4.err
python examples/polar_sweep.py --steps 4000 --alpha-min 0 --alpha-max 0 --out $O/t12_polar1.csv > $O/t12_polar1_single.json 2>> $O/t12_polar4.err
echo done
