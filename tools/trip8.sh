#!/bin/bash
# GPU trip 8 (2 GPUs): multi-GPU tests after the kernel-preload fix; weak-scaling diagnosis
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_slabs.py -q -v > $O/t8_pytest_multi.log 2>&1; echo "rc=$?" >> $O/t8_pytest_multi.log
timeout 120 python tools/diag_multi.py 8 80 0 0 0 > $O/t8_diag.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
W="bench.py --steps 200 --warmup 20 --scaling weak --no-e2e --no-cpu-baseline"
python $W > $O/t8_w1.json 2> $O/t8_w1.err
AEROLAB_LBM_FAKE_HALO=1 python $W > $O/t8_w1_fake.json 2> $O/t8_w1_fake.err
AEROLAB_LBM_TRACE=100 $TR --nproc-per-node 2 --master-port 29611 $W --gpus 2 > $O/t8_w2.json 2> $O/t8_w2.err
AEROLAB_LBM_S2_GENERATIONS=24 $TR --nproc-per-node 2 --master-port 29612 $W --gpus 2 > $O/t8_w2_gen24.json 2> $O/t8_w2_gen24.err
AEROLAB_LBM_AUX_PRIO=0 $TR --nproc-per-node 2 --master-port 29613 $W --gpus 2 > $O/t8_w2_prio0.json 2> $O/t8_w2_prio0.err
AEROLAB_LBM_DOUBLE=0 $TR --nproc-per-node 2 --master-port 29614 $W --gpus 2 > $O/t8_w2_single.json 2> $O/t8_w2_single.err
AEROLAB_LBM_DOUBLE=0 python $W > $O/t8_w1_single.json 2> $O/t8_w1_single.err
echo done
