#!/bin/bash
# A/B of library variants on one box, models built once per session.
# usage: scripts/abx.sh <tag> "<lib suffixes ('' = libr3dgpu.so)>" "<cfg deg n;...>" [reps]
tag=$1; out=gpurun_out/$tag; mkdir -p $out
IFS=';' read -ra WL <<< "$3"
export R3D_REPS=${4:-3}
for w in "${WL[@]}"; do
  for v in $2; do
    [ "$v" = "default" ] && v=""
    R3D_LIBRARY=$PWD/radiative3d_b200/libr3dgpu$v.so timeout 600 python scripts/profile_target.py $w 2>&1 | tail -${ABX_TAIL:-1} | sed "s/^/[lib$v] /" | tee -a $out/ab.log
  done
done
