#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
( time python -m pytest tests -m gpu -q --maxfail=12 --durations=8 ) > $O/r02b_pytest.log 2>&1
tail -40 $O/r02b_pytest.log
echo "== cornell principled 128 spp"
BENCH_ARGS="--workload cornell --spp 128" tools/variants.sh run base mb2 mb4 a24
echo "== terrain 64 spp"
BENCH_ARGS="--spp 64" tools/variants.sh run base mb2 mb4 a24
echo "== cube 64 spp (multiscatter principled)"
BENCH_ARGS="--workload cube --spp 64" tools/variants.sh run base
