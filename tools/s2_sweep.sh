#!/bin/bash
# A/B sweep of the fused two-step kernel: variants/lib_s2_*.so are builds with other -DALB_S2_* shapes
B="python bench.py --workload configs[3] --steps 41 --warmup 5 --no-cpu-baseline --no-e2e"
out=gpurun_out/r2_s2_sweep.log
: > $out
for hs in 64 128 256; do
  echo "default hs=$hs" >> $out
  AEROLAB_LBM_S2_HS=$hs $B 2>&1 | tail -1 | cut -c1-120 >> $out
done
for lib in variants/lib_s2_*.so; do
  for hs in 128; do
    echo "$lib hs=$hs" >> $out
    AEROLAB_LBM_LIB=$PWD/$lib AEROLAB_LBM_S2_HS=$hs $B 2>&1 | tail -1 | cut -c1-120 >> $out
  done
done
cat $out
