#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_render_gpu.py -m gpu -q -k "block_shapes or ambient or principled or sampling" 2>&1 | tail -3
for m in 2 -1; do
  echo "== lean multiscatter, shade_wide=$m"; SHADE_WIDE=$m timeout 300 python tools/shade_cap_ab.py 2>&1 | tail -5
done
for m in 2 -1; do
  echo "== full kernels, shade_wide=$m"; SHADE_WIDE=$m timeout 300 python tools/shade_cap_ext_ab.py 2>&1 | tail -4
done
