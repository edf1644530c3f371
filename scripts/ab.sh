#!/bin/bash
# Scratch: A/B of library variants on one box.  usage: scripts/ab.sh <tag> "<lib suffixes>" "<cfg deg n;...>"
# e.g. scripts/ab.sh s5 "_e0 '' _s1" "halfspace_nearsrc50 9 1e8;lopnor 6 1e7"     ('' = libr3dgpu.so)
tag=$1; out=gpurun_out/$tag; mkdir -p $out
IFS=';' read -ra WL <<< "$3"
for rep in 1 2; do
for w in "${WL[@]}"; do
  for v in $2; do
    [ "$v" = "''" ] && v=""
    R3D_LIBRARY=$PWD/radiative3d_b200/libr3dgpu$v.so timeout 300 python scripts/profile_target.py $w 2>&1 | tail -1 | sed "s/^/[lib$v] /" | tee -a $out/ab.log
  done
done
done
