#!/bin/bash
# shared-memory top of the traversal stack (SMEM_STACK = 4 / 8 entries) vs local memory, A/B
cd "$(dirname "$0")/.."
timeout 300 env B200_CYCLES_LIB=$PWD/raytracingproject_b200/_build/lib_ss8.so python -m pytest tests/test_traversal_gpu.py tests/test_golden_gpu.py -m gpu -q -x 2>&1 | tail -2
echo "== terrain 64 spp"
BENCH_ARGS="--spp 64" tools/variants.sh run base ss4 ss8 base
echo "== instanced 16 spp"
BENCH_ARGS="--workload instanced --spp 16" tools/variants.sh run base ss4 ss8 base
echo "== cornell 128 spp"
BENCH_ARGS="--workload cornell --spp 128" tools/variants.sh run base ss4 ss8
