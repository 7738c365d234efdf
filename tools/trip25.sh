#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python tools/sweep_double.py 512x256 768x384 1024x512 1280x640 1536x768 2048x512 2048x1024 2000x1000 3072x1536 4096x1024 4000x2000 4096x2048 1024x4096 > $O/t25_sweep.log 2>&1
echo done
