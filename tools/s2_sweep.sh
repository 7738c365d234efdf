#!/bin/bash
# A/B sweep of the fused two-step kernel: variants/lib_s2_*.so are builds with other -DALB_S2_* settings
B="python bench.py --workload configs[3] --steps 41 --warmup 5 --no-cpu-baseline --no-e2e"
out=gpurun_out/r2_s2_sweep.log
: > $out
echo "default" >> $out
$B 2>&1 | tail -1 | cut -c1-120 >> $out
for lib in variants/lib_s2_*.so; do
  echo "$lib" >> $out
  AEROLAB_LBM_LIB=$PWD/$lib $B 2>&1 | tail -1 | cut -c1-120 >> $out
done
cat $out
