#!/bin/bash
cd "$(dirname "$0")/.."
echo "== cornell principled 128 spp"
BENCH_ARGS="--workload cornell --spp 128" tools/variants.sh run old base loc2 loc3 fast fastloc2
echo "== terrain 64 spp"
BENCH_ARGS="--spp 64" tools/variants.sh run old base loc2 fast fastloc2
