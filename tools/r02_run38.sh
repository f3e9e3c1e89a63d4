#!/bin/bash
cd "$(dirname "$0")/.."
echo "== terrain 64 spp (base = built with -split-compile 0)"
BENCH_ARGS="--spp 64" tools/variants.sh run base nosplit base nosplit
echo "== instanced 16 spp"
BENCH_ARGS="--workload instanced --spp 16" tools/variants.sh run base nosplit
echo "== cornell 64 spp"
BENCH_ARGS="--workload cornell --spp 64" tools/variants.sh run base nosplit
