#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/t6_diag.log
: > $O
for i in 1 2 3; do echo "== run $i: two_device" >> $O; timeout 200 python -m pytest tests/test_gpu_multi.py -q -k two_device 2>&1 | tail -4 >> $O; done
echo "== only 1400" >> $O; timeout 200 python -m pytest tests/test_gpu_multi.py -q -k "1400" 2>&1 | tail -4 >> $O
echo "== DEFER=0 two_device" >> $O; AEROLAB_LBM_DEFER_SIGNAL=0 timeout 200 python -m pytest tests/test_gpu_multi.py -q -k two_device 2>&1 | tail -4 >> $O
echo "== PRIO=0 two_device" >> $O; AEROLAB_LBM_AUX_PRIO=0 timeout 200 python -m pytest tests/test_gpu_multi.py -q -k two_device 2>&1 | tail -4 >> $O
echo "== MAXCONN=8 two_device" >> $O; CUDA_DEVICE_MAX_CONNECTIONS=8 timeout 200 python -m pytest tests/test_gpu_multi.py -q -k two_device 2>&1 | tail -4 >> $O
echo done
