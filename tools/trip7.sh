#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/t7_diag.log
: > $O
for args in "8 80 1 0 0" "8 80 1 0 1" "8 80 0 0 0" "2 80 1 0 0" "8 80 1 -1 0"; do
  echo "== $args" >> $O; timeout 120 python tools/diag_multi.py $args >> $O 2>&1
done
echo "== NO_GRAPH: 8 80 1 0 0" >> $O; AEROLAB_LBM_NO_GRAPH=1 timeout 120 python tools/diag_multi.py 8 80 1 0 0 >> $O 2>&1
echo "== TRACE: 8 80 1 0 0" >> $O; AEROLAB_LBM_TRACE=8 timeout 120 python tools/diag_multi.py 8 80 1 0 0 >> $O 2>&1
echo done
