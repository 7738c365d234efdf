#!/bin/bash
# full 1-GPU check of the tree with inlet/outlet columns in march2_kernel + ncu evidence for it
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > $O/t26_pytest.log 2>&1; echo "rc=$?" >> $O/t26_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/t26_smoke.log 2>&1; echo "rc=$?" >> $O/t26_smoke.log
timeout 300 python tools/sweep_double.py 2000x1000 1400x1400 2048x1024 > $O/t26_sweep.log 2>&1
python bench.py > $O/t26_line.json 2> $O/t26_line.err
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-weak"
$B > $O/t26_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:march2 -s 3 -c 1 -f -o $O/r2h_march_edges_c3 $B > $O/t26_ncu.log 2>&1
B2="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-weak"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2h_launches_c3_raw.csv $B2 > $O/t26_ncu2.log 2>&1
echo done
