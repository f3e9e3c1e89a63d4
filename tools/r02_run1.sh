#!/bin/bash
# round 2, first check of the new bench / tests on the GPU box
cd "$(dirname "$0")/.."
O=gpurun_out
( time python -m pytest tests -m gpu -x -q --durations=15 ) > $O/r02a_pytest.log 2>&1
tail -5 $O/r02a_pytest.log
python bench.py --steps 5 --warmup 3 > $O/r02a_bench_n1.json 2> $O/r02a_bench_n1.err
tail -c 600 $O/r02a_bench_n1.err
B="--no-cpu-baseline --no-e2e --configs none"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_intersect_closest --launch-count 2 \
  -o $O/r02a_closest_terrain -f python bench.py --spp 16 --steps 1 --warmup 0 $B > $O/ncu_a1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_intersect_closest --launch-count 3 \
  -o $O/r02a_closest_instanced -f python bench.py --workload instanced --spp 4 --steps 1 --warmup 0 $B > $O/ncu_a2.log 2>&1
ls -la $O | tail -8
