#!/bin/bash
# denoising data passes against the reference + the register-cap A/B of k_shade_surface
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_render_gpu.py -m gpu -q -s -k "denoising or passes" 2>&1 | grep -E "^denoising|passed|failed|Error|assert" | cut -c1-200 | tail -80
echo "== cornell multiscatter 128 spp"
BENCH_ARGS="--workload cornell --spp 128" tools/variants.sh run base mb3 mb4 base
echo "== cube"
BENCH_ARGS="--workload cube" tools/variants.sh run base mb3 mb4
echo "== terrain 64 spp"
BENCH_ARGS="--spp 64" tools/variants.sh run base mb3
