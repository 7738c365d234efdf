#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
AEROLAB_LBM_LIB=$PWD/variants/v_w12.so python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/t18_c3_w12.json 2> $O/t18_c3_w12.err
python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/t18_c3.json 2> $O/t18_c3.err
AEROLAB_LBM_LIB=$PWD/variants/v_w12.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:march2 -s 3 -c 1 -o $O/r2f_march_w12_c3 \
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > $O/t18_ncu.log 2>&1
echo done
