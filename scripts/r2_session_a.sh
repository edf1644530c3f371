#!/bin/bash
# round 2, session a: GPU tests of the starting build, then ncu --set full of the curved-ray kernels at TOA degree 9
out=gpurun_out/r2a; mkdir -p $out
nproc > $out/nproc.log; nvidia-smi -L >> $out/nproc.log
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; tail -3 $out/pytest_gpu.log
for cfg in spherical:2e6 crustpinch:1e7 lopnor:1e7; do
  c=${cfg%%:*}; n=${cfg##*:}
  timeout 600 python scripts/profile_target.py $c 9 $n > $out/plain_$c.log 2>&1; cat $out/plain_$c.log
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:propagate_kernel -s 1 -c 1 -f -o $out/prof_$c \
    python scripts/profile_target.py $c 9 $n > $out/ncu_$c.log 2>&1; tail -1 $out/ncu_$c.log
done
