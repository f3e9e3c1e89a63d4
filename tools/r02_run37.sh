#!/bin/bash
cd "$(dirname "$0")/.."
echo "== terrain 64 spp"
BENCH_ARGS="--spp 64" tools/variants.sh run base ggx1 ggx2 base
echo "== instanced 16 spp"
BENCH_ARGS="--workload instanced --spp 16" tools/variants.sh run base ggx1 ggx2
