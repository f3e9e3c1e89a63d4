#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests/test_render_gpu.py -m gpu -q -s -k "passes or invisible or clamp or transparent" 2>&1 | grep -E "passes_transparent|clamp_after|passed|failed|Error" | cut -c1-200 | tail -24
echo "== cornell multiscatter 128 spp"
BENCH_ARGS="--workload cornell --spp 128" tools/variants.sh run base me ls both base
echo "== cornell GGX 128 spp"
BENCH_ARGS="--workload cornell --spp 128 --distribution GGX" tools/variants.sh run base me ls both
echo "== cube"
BENCH_ARGS="--workload cube" tools/variants.sh run base me ls both
echo "== terrain 64 spp"
BENCH_ARGS="--spp 64" tools/variants.sh run base me ls both
