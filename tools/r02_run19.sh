#!/bin/bash
# ray-tile staging: 1-D bulk async copy + mbarrier (shipped) vs LDGSTS cp.async (variant), A/B
cd "$(dirname "$0")/.."
echo "== terrain 64 spp"
BENCH_ARGS="--spp 64" tools/variants.sh run base ldgsts base ldgsts
echo "== instanced 16 spp"
BENCH_ARGS="--workload instanced --spp 16" tools/variants.sh run base ldgsts base ldgsts
echo "== cornell 128 spp"
BENCH_ARGS="--workload cornell --spp 128" tools/variants.sh run base ldgsts
echo "== cube"
BENCH_ARGS="--workload cube" tools/variants.sh run base ldgsts
