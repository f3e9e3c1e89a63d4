#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_render_gpu.py -m gpu -q -k "register_budgets or ambient or sampling or adaptive" 2>&1 | tail -4
for m in 0 1 -1; do
  echo "== shipped lib shade_dense=$m"; SHADE_DENSE=$m timeout 300 python tools/shade_cap_ext_ab.py 2>&1 | tail -4
done
