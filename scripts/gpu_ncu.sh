#!/bin/bash
# ncu --set full capture of the propagate kernel on the bench workload.  usage: scripts/gpu_ncu.sh <tag> [n phonons] [config] [toa degree]
tag=${1:-x}; n=${2:-2e7}; cfg=${3:-halfspace_nearsrc50}; deg=${4:-9}
out=gpurun_out/$tag; mkdir -p $out
timeout 300 python scripts/profile_target.py $cfg $deg $n > $out/plain_target.log 2>&1; cat $out/plain_target.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:propagate_kernel -s 1 -c 1 -f -o $out/prof_propagate \
  python scripts/profile_target.py $cfg $deg $n > $out/ncu_propagate.log 2>&1; tail -1 $out/ncu_propagate.log
