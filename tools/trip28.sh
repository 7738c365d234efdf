#!/bin/bash
# 1 GPU: four configs[4] cases (2048x1024, 20,000 steps each) sharing one GPU -- the per-GPU load of the
# 31-case polar on 8 GPUs (which shards cases over ranks with no communication)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
python examples/polar_sweep.py --steps 20000 --alpha-min 2 --alpha-max 5 --out $O/r2h_polar4_naca0012_2048x1024_20000steps.csv > $O/t28_polar4.json 2> $O/t28_polar4.err
echo done
