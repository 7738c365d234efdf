#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
python tools/trace_case.py 2048 1024 > $O/t23_trace_2048.log 2>&1
python tools/trace_case.py 4096 2048 naca4412 10 > $O/t23_trace_4096.log 2>&1
AEROLAB_LBM_NO_GRAPH= python tools/trace_case.py 2048 1024 > $O/t23_trace_2048_b.log 2>&1
echo done
