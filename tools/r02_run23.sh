#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_render_gpu.py -m gpu -q -s -k "denoising_transparent" 2>&1 | grep -E "shadow_a|shadow_b|passed|failed|Error|assert" | cut -c1-200 | tail -30
for l in base mb3 mb4; do
  lib=$PWD/raytracingproject_b200/_build/lib_$l.so
  [ $l = base ] && lib=$PWD/raytracingproject_b200/libb200cycles.so
  echo "== $l"; B200_CYCLES_LIB=$lib timeout 300 python tools/shade_cap_ab.py 2>&1 | tail -5
done
