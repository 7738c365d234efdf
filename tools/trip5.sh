#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/t5_diag.log
: > $O
for args in "8 80 1" "2 80 1" "4 80 1" "8 16 1" "8 80 0" "16 80 1"; do
  echo "== default: $args" >> $O; timeout 120 python tools/diag_multi.py $args >> $O 2>&1
done
echo "== DEFER=0: 8 80 1" >> $O; AEROLAB_LBM_DEFER_SIGNAL=0 timeout 120 python tools/diag_multi.py 8 80 1 >> $O 2>&1
echo "== PRIO=0: 8 80 1" >> $O; AEROLAB_LBM_AUX_PRIO=0 timeout 120 python tools/diag_multi.py 8 80 1 >> $O 2>&1
echo "== PRIO=0 DEFER=0: 8 80 1" >> $O; AEROLAB_LBM_DEFER_SIGNAL=0 AEROLAB_LBM_AUX_PRIO=0 timeout 120 python tools/diag_multi.py 8 80 1 >> $O 2>&1
echo "== MAXCONN=32: 8 80 1" >> $O; CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 120 python tools/diag_multi.py 8 80 1 >> $O 2>&1
echo done
