#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
( time python -m pytest tests -m gpu -q -s --maxfail=12 ) > $O/r02d_pytest.log 2>&1
tail -25 $O/r02d_pytest.log | cut -c1-300
echo "== cornell principled 128 spp"
BENCH_ARGS="--workload cornell --spp 128" tools/variants.sh run old fastloc2 base
echo "== terrain 64 spp"
BENCH_ARGS="--spp 64" tools/variants.sh run old base
echo "== instanced 4K 16 spp"
BENCH_ARGS="--workload instanced --spp 16" tools/variants.sh run old base
B="--no-cpu-baseline --no-e2e --configs none"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --workload cornell --spp 32 --steps 1 --warmup 1 $B > $O/ncu_ll.log 2>&1
