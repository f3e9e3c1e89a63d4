#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_render_gpu.py -m gpu -q -s -k "register_budgets or denoising or principled" 2>&1 | grep -E "dense|passed|failed|Error|assert" | cut -c1-200 | tail -30
for w in cube cornell; do
for m in -1 0 1; do
  echo "== $w shade_dense=$m"
  python bench.py --steps 3 --warmup 3 --workload $w --spp 64 --no-cpu-baseline --no-e2e --configs none --opt shade_dense=$m 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('Mrays/s %.0f ms/step %.1f' % (d['value'], d['ms_per_step']), d['control']['share'], 'dense', d['control'].get('shade_dense'))
"
done
done
