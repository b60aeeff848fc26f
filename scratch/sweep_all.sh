#!/bin/bash
# each geometry in its own process with a short timeout (a hang costs seconds, not the GPU box)
W=${1:-cityscapes_1024x2048_b8_n100}
for c in 2x8x2 2x8x3 4x4x3 4x4x4 2x4x4 2x4x6 4x2x6; do
  timeout 45 python scratch/sweep_dense.py $W v2,$c > gpurun_out/sw_$c.log 2>&1; rc=$?
  echo "$c rc=$rc $(tail -1 gpurun_out/sw_$c.log | cut -c1-200)"
done
