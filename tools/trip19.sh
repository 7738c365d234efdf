#!/bin/bash
# GPU trip 19 (2 GPUs): final code -- whole suite incl. multi-GPU tests, smoke, N=1 and N=2 lines (strong, weak), configs[1], [2], [4]
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > $O/t19_pytest.log 2>&1; echo "rc=$?" >> $O/t19_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/t19_smoke.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python bench.py > $O/t19_s1.json 2> $O/t19_s1.err
python bench.py --impl reference > $O/t19_ref.json 2> $O/t19_ref.err
$TR --nproc-per-node 2 --master-port 29801 bench.py --gpus 2 > $O/t19_s2.json 2> $O/t19_s2.err
python bench.py --scaling weak --no-cpu-baseline > $O/t19_w1.json 2> $O/t19_w1.err
$TR --nproc-per-node 2 --master-port 29802 bench.py --gpus 2 --scaling weak > $O/t19_w2.json 2> $O/t19_w2.err
python bench.py --workload "configs[1]" --steps 20000 --warmup 2000 --no-cpu-baseline > $O/t19_c1.json 2> $O/t19_c1.err
python bench.py --workload "configs[2]" --steps 2000 --warmup 200 --no-cpu-baseline > $O/t19_c2.json 2> $O/t19_c2.err
python bench.py --workload "configs[4]-case" --steps 4000 --warmup 400 --no-cpu-baseline > $O/t19_c4.json 2> $O/t19_c4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2f_launches_c3_raw.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/t19_ncu2.log 2>&1
echo done
