#!/bin/bash
# One GPU-box session: parity tests, the bench line, the reference arm, the ncu launch list and one full
# capture per kernel.  Everything is written under gpurun_out/<tag>/.   usage: scripts/gpu_session.sh <tag> [what...]
tag=${1:-r1}; shift
what=${*:-"tests bench ref launches ncu"}
out=gpurun_out/$tag; mkdir -p $out
small="--steps 1 --warmup 3 --per-gpu 4000000 --no-cpu-baseline --e2e-steps 0"
for w in $what; do case $w in
  tests)    timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; tail -3 $out/pytest_gpu.log;;
  smoke)    timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $out/smoke.log 2>&1; tail -2 $out/smoke.log;;
  bench)    timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; cat $out/bench.json; tail -3 $out/bench.err;;
  ref)      timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; cat $out/bench_ref.json;;
  launches) timeout 600 python bench.py $small > $out/plain_small.json 2>&1 && \
            timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
              python bench.py $small > $out/ncu_launches.log 2>&1; tail -1 $out/ncu_launches.log | cut -c1-200;;
  ncu)      scripts/gpu_ncu.sh $tag 2e7;;
  create)   timeout 300 python scripts/time_create.py > $out/time_create.log 2>&1; cat $out/time_create.log;;
esac; done
