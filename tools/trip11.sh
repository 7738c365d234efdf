#!/bin/bash
# GPU trip 11 (8 GPUs): strong / weak scaling lines at N = 4, 8 with the final code, the configs[4] polar,
# a field image of configs[3] from 8 slabs, the multi-process bitwise tests on 4 ranks
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nvidia-smi -L > $O/t11_gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29701 bench.py --gpus 8 --steps 200 --warmup 20 > $O/t11_s8.json 2> $O/t11_s8.err
$TR --nproc-per-node 4 --master-port 29702 bench.py --gpus 4 --steps 200 --warmup 20 > $O/t11_s4.json 2> $O/t11_s4.err
$TR --nproc-per-node 8 --master-port 29703 bench.py --gpus 8 --steps 200 --warmup 20 --scaling weak > $O/t11_w8.json 2> $O/t11_w8.err
$TR --nproc-per-node 4 --master-port 29704 bench.py --gpus 4 --steps 200 --warmup 20 --scaling weak > $O/t11_w4.json 2> $O/t11_w4.err
$TR --nproc-per-node 8 --master-port 29705 examples/polar_sweep.py --steps 20000 --out $O/r2_polar_naca0012_2048x1024_20000steps.csv > $O/t11_polar.json 2> $O/t11_polar.err
$TR --nproc-per-node 8 --master-port 29706 examples/slab_field_png.py --mode vort --steps 3000 --stride 8 --out $O/r2_configs3_vort_8gpu.png > $O/t11_png.log 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -q > $O/t11_pytest_multi.log 2>&1; echo "rc=$?" >> $O/t11_pytest_multi.log
echo done
