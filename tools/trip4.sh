#!/bin/bash
# GPU trip 4 of round 2 (2 GPUs): multi-GPU bitwise tests with the final protocol, N=1/2 scaling lines
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nvidia-smi -L > $O/t4_gpus.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_slabs.py -q -v > $O/t4_pytest_multi.log 2>&1; echo "rc=$?" >> $O/t4_pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/t4_n1.json 2> $O/t4_n1.err
$TR --nproc-per-node 2 --master-port 29601 bench.py --gpus 2 --steps 200 --warmup 20 > $O/t4_n2.json 2> $O/t4_n2.err
$TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --steps 200 --warmup 20 --scaling weak > $O/t4_n2_weak.json 2> $O/t4_n2_weak.err
python bench.py --steps 200 --warmup 20 --scaling weak --no-cpu-baseline > $O/t4_n1_weak.json 2> $O/t4_n1_weak.err
AEROLAB_LBM_AUX_PRIO=0 $TR --nproc-per-node 2 --master-port 29603 bench.py --gpus 2 --steps 200 --warmup 20 --no-e2e > $O/t4_n2_prio0.json 2> $O/t4_n2_prio0.err
AEROLAB_LBM_S2_GENERATIONS=24 $TR --nproc-per-node 2 --master-port 29604 bench.py --gpus 2 --steps 200 --warmup 20 --no-e2e > $O/t4_n2_gen24.json 2> $O/t4_n2_gen24.err
AEROLAB_LBM_S2_GENERATIONS=24 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e > $O/t4_n1_gen24.json 2> $O/t4_n1_gen24.err
AEROLAB_LBM_S2_GENERATIONS=48 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e > $O/t4_n1_gen48.json 2> $O/t4_n1_gen48.err
echo done
