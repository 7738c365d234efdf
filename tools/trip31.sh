#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
AEROLAB_LBM_TRACE_DESTROY=1 python tools/case_phases.py > $O/t31_phases.log 2>&1
timeout 300 python -m pytest tests/test_gpu_band.py tests/test_gpu_parity.py -m gpu -q -x > $O/t31_pytest.log 2>&1; echo "rc=$?" >> $O/t31_pytest.log
echo done
