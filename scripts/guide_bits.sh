#!/bin/bash
# Scratch: kernel time and r3d_create time against the size of the guide tables.  usage: scripts/guide_bits.sh <tag> "<bits...>" "<cfg deg n;...>"
tag=$1; out=gpurun_out/$tag; mkdir -p $out
IFS=';' read -ra WL <<< "$3"
for rep in 1 2; do for w in "${WL[@]}"; do for b in $2; do
  R3D_GUIDE_BITS=$b timeout 300 python scripts/profile_target.py $w 2>&1 | tail -1 | sed "s/^/[bits $b] /" | tee -a $out/bits.log
done; done; done
for b in $2; do echo "== bits $b" | tee -a $out/create.log; R3D_GUIDE_BITS=$b R3D_TIMING=1 timeout 300 python scripts/time_create.py 2>&1 | grep -E "guide|create" | tail -4 | tee -a $out/create.log; done
