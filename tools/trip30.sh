#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
python tools/case_phases.py > $O/t30_phases.log 2>&1
python bench.py --workload "configs[2]" --steps 2000 --warmup 200 --no-cpu-baseline > $O/t30_c2.json 2> $O/t30_c2.err
python bench.py --workload "configs[4]-case" --steps 4000 --warmup 400 --no-cpu-baseline > $O/t30_c4.json 2> $O/t30_c4.err
python bench.py --workload "configs[1]" --steps 20000 --warmup 2000 --no-cpu-baseline > $O/t30_c1.json 2> $O/t30_c1.err
echo done
