#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_band.py tests/test_reference_pins.py -m gpu -q -x > $O/t21_band.log 2>&1; echo "rc=$?" >> $O/t21_band.log
python bench.py --workload "configs[1]" --steps 20000 --warmup 2000 --no-cpu-baseline > $O/t21_c1.json 2> $O/t21_c1.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:band_lattice -s 1 -c 1 -o $O/r2g_band_c1 \
  python bench.py --workload "configs[1]" --steps 2000 --warmup 200 --no-cpu-baseline --no-e2e > $O/t21_ncu.log 2>&1
echo done
