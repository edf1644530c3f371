#!/bin/bash
# round 2: full GPU tests of the build with the slot diet / convergent catch / reciprocal + rsqrt diet, slot-count A/B, ncu of the halfspace kernel
out=gpurun_out/r2h; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; tail -3 $out/pytest_gpu.log
scripts/env_ab.sh r2h "halfspace_nearsrc50 9 1.25e8" ";R3D_SLOTS_PER_BLOCK=1152;R3D_SLOTS_PER_BLOCK=1248;R3D_SLOTS_PER_BLOCK=1344;R3D_SMEM_KB=164;R3D_SMEM_KB=228" > /dev/null 2>&1
scripts/env_ab.sh r2h "spherical 9 2e6" ";R3D_SLOTS_PER_BLOCK=1024;R3D_SLOTS_PER_BLOCK=1280;R3D_THREADS=384" > /dev/null 2>&1
scripts/env_ab.sh r2h "crustpinch 9 1e7" ";R3D_SLOTS_PER_BLOCK=1024;R3D_SLOTS_PER_BLOCK=1280;R3D_THREADS=384" > /dev/null 2>&1
cat $out/env_ab.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:propagate_kernel -s 1 -c 1 -f -o $out/prof_halfspace_nearsrc50 python scripts/profile_target.py halfspace_nearsrc50 9 2e7 > $out/ncu_hs.log 2>&1; tail -1 $out/ncu_hs.log
