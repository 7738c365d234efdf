#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > $O/t17_pytest.log 2>&1; echo "rc=$?" >> $O/t17_pytest.log
python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/t17_c3.json 2> $O/t17_c3.err
AEROLAB_LBM_LIB=$PWD/variants/v_diag16.so python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/t17_c3_diag16.json 2> $O/t17_c3_diag16.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:march2 -c 12 --csv --log-file $O/t17_march_times.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/t17_ncu.log 2>&1
echo done
