#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > $O/t14_pytest.log 2>&1; echo "rc=$?" >> $O/t14_pytest.log
python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/t14_c3.json 2> $O/t14_c3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:march2 -c 12 --csv --log-file $O/t14_march_times.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/t14_ncu.log 2>&1
echo done
