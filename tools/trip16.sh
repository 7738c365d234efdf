#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_double.py tests/test_gpu_slabs.py tests/test_reference_pins.py tests/test_gpu_large.py -m gpu -q -x > $O/t16_pytest.log 2>&1; echo "rc=$?" >> $O/t16_pytest.log
python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/t16_c3.json 2> $O/t16_c3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:march2 -c 12 --csv --log-file $O/t16_march_times.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/t16_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march2 -s 0 -c 1 -o $O/r2e_march_diag_c3 \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > $O/t16_ncu2.log 2>&1
echo done
