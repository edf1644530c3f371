#!/bin/bash
# round 2: chunk clocks of the five workloads + ncu --set full of the tetrahedral kernel (build with the arc diet)
out=gpurun_out/r2l; mkdir -p $out
export R3D_TIMING=1 ABX_TAIL=7
scripts/abx.sh r2l "_clocks" "halfspace_nearsrc50 9 1.25e8;crustpinch 9 1e7;lopnor 9 2e7;spherical 9 2e6" 1 > /dev/null 2>&1
grep -v r3d_create $out/ab.log
unset R3D_TIMING
timeout 600 ncu --set full --clock-control none --import-source on -k regex:propagate_kernel -s 1 -c 1 -f -o $out/prof_crustpinch python scripts/profile_target.py crustpinch 9 1e7 > $out/ncu_crustpinch.log 2>&1; tail -1 $out/ncu_crustpinch.log
python scripts/dropin_e2e.py lopnor 1e8 2>&1 | tee $out/dropin.log
