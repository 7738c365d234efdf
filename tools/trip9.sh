#!/bin/bash
# GPU trip 9 (1 GPU): whole suite with the band kernel, small-lattice timing, final march defaults, ncu
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > $O/t9_pytest.log 2>&1; echo "rc=$?" >> $O/t9_pytest.log
python bench.py --workload "configs[1]" --steps 20000 --warmup 2000 --no-cpu-baseline > $O/t9_c1_band.json 2> $O/t9_c1_band.err
AEROLAB_LBM_BAND=0 python bench.py --workload "configs[1]" --steps 20000 --warmup 2000 --no-cpu-baseline > $O/t9_c1_grid.json 2> $O/t9_c1_grid.err
python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e > $O/t9_c3.json 2> $O/t9_c3.err
python bench.py --workload "configs[2]" --steps 2000 --warmup 200 --no-cpu-baseline > $O/t9_c2.json 2> $O/t9_c2.err
python bench.py --workload "configs[4]-case" --steps 4000 --warmup 400 --no-cpu-baseline > $O/t9_c4.json 2> $O/t9_c4.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/t9_smoke.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march2 -s 3 -c 1 -o $O/r2c_march_c3 \
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > $O/t9_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2c_launches_c3_raw.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/t9_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:band_lattice -s 1 -c 1 -o $O/r2c_band_c1 \
  python bench.py --workload "configs[1]" --steps 2000 --warmup 200 --no-cpu-baseline --no-e2e > $O/t9_ncu3.log 2>&1
echo done
