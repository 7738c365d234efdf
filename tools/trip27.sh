#!/bin/bash
# 2 GPUs: multi-GPU tests and N=2 lines with the inlet/outlet columns in march2_kernel
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q > $O/t27_pytest_multi.log 2>&1; echo "rc=$?" >> $O/t27_pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 2 --master-port 29801 bench.py --gpus 2 --no-weak > $O/t27_s2.json 2> $O/t27_s2.err
$TR --nproc-per-node 2 --master-port 29802 bench.py --gpus 2 --scaling weak > $O/t27_w2.json 2> $O/t27_w2.err
echo done
