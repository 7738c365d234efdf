#!/bin/bash
# segment-height sweep (AEROLAB_LBM_S2_HS) for the default build and one variant library
B="python bench.py --workload configs[3] --steps 41 --warmup 5 --no-cpu-baseline --no-e2e"
out=gpurun_out/r2_s2_hs.log
: > $out
for lib in "" variants/lib_s2_ld_rb3_k4_pf0.so; do
for hs in 128 120 136 100 200 250; do
  echo "lib=$lib hs=$hs" >> $out
  if [ -z "$lib" ]; then AEROLAB_LBM_S2_HS=$hs $B 2>&1 | tail -1 | cut -c1-120 >> $out
  else AEROLAB_LBM_LIB=$PWD/$lib AEROLAB_LBM_S2_HS=$hs $B 2>&1 | tail -1 | cut -c1-120 >> $out; fi
done; done
cat $out
