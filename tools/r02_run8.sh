#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests/test_render_gpu.py tests/test_full_gates_gpu.py -m gpu -q --maxfail=5 2>&1 | tail -3
echo "== cube multiscatter 64 spp"
BENCH_ARGS="--workload cube" tools/variants.sh run lm0 base
echo "== cornell multiscatter 128 spp"
BENCH_ARGS="--workload cornell --spp 128" tools/variants.sh run lm0 base
echo "== cornell GGX 128 spp"
BENCH_ARGS="--workload cornell --spp 128 --distribution GGX" tools/variants.sh run lm0 base
echo "== terrain 64 spp"
BENCH_ARGS="--spp 64" tools/variants.sh run lm0 base
echo "== instanced (100k-tri BLAS) 4K 16 spp"
BENCH_ARGS="--workload instanced --spp 16" tools/variants.sh run base
