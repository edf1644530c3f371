#!/bin/bash
# Scratch: A/B of launch-geometry environment settings on one box.  usage: scripts/env_ab.sh <tag> "<cfg deg n>" "ENV1=.. ENV2=..;ENV=..;..."   (an empty entry = defaults)
tag=$1; out=gpurun_out/$tag; mkdir -p $out
IFS=';' read -ra EV <<< "$3"
for rep in 1 2; do for e in "${EV[@]}"; do
  env $e timeout 300 python scripts/profile_target.py $2 2>&1 | tail -1 | sed "s/^/[${e:-defaults}] /" | tee -a $out/env_ab.log
done; done
