#!/bin/bash
# round 2, measurement session on one B200: parity tests, ncu launch list of the bench command, one `ncu --set full` capture
# of the propagate kernel per BASELINE workload (summarised on the box: five reports exceed what gpurun brings back), the five
# bench lines and the reference arm.   usage: scripts/r2_session_final.sh [tag]
tag=${1:-r2}; out=gpurun_out/$tag; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; tail -3 $out/pytest_gpu.log
small="--steps 1 --warmup 3 --per-gpu 4000000 --no-cpu-baseline --e2e-steps 0"
timeout 600 python bench.py $small > $out/plain_small.json 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py $small > $out/ncu_launches.log 2>&1
specs=""
# (layered models: the kernel build the pilot launches pick is forced here, so that launch 1 is the whole measured job)
for w in "halfspace_nearsrc50 2e7 1" "halfspace 2e7 1" "crustpinch 1e7 -1" "lopnor 1e7 0" "spherical 2e6 -1"; do
  set -- $w; c=$1; n=$2
  R3D_CYL_WIDE=$3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:propagate_kernel -s 1 -c 1 -f -o /tmp/prof_$c \
    python scripts/profile_target.py $c 9 $n > $out/ncu_$c.log 2>&1
  tail -1 $out/ncu_$c.log
  python scripts/ncu_summary.py /tmp/prof_$c.ncu-rep > $out/summary_$c.md 2>&1
  python scripts/ncu_stalls.py /tmp/prof_$c.ncu-rep 1000 > $out/stalls_$c.txt 2>&1
  specs="$specs $c=/tmp/prof_$c.ncu-rep:$out/ncu_$c.log"
done
R3D_COUNTERS_OUT=$PWD/$out/kernel_counters.json python scripts/ncu_counters.py ${tag} $specs > $out/counters.log 2>&1
cp $out/kernel_counters.json profiles/kernel_counters.json
cp /tmp/prof_halfspace_nearsrc50.ncu-rep /tmp/prof_spherical.ncu-rep $out/ 2>/dev/null
for c in halfspace_nearsrc50 halfspace crustpinch lopnor spherical; do
  timeout 1200 python bench.py --workload $c > $out/bench_$c.json 2> $out/bench_$c.err; cut -c1-400 $out/bench_$c.json; tail -2 $out/bench_$c.err
done
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err; cat $out/bench_reference.json
