#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
for v in ds fm exact; do
  B200_CYCLES_LIB=$PWD/raytracingproject_b200/_build/lib_$v.so python -m pytest tests/test_render_gpu.py tests/test_svm_nodes_gpu.py tests/test_full_gates_gpu.py tests/test_golden_gpu.py tests/test_traversal_gpu.py -m gpu -q -s --maxfail=30 > $O/r02e_pytest_$v.log 2>&1
  echo "== $v"; tail -6 $O/r02e_pytest_$v.log | cut -c1-200
done
echo "== cornell principled 128 spp"
BENCH_ARGS="--workload cornell --spp 128" tools/variants.sh run old exact ds fm base
echo "== terrain 64 spp"
BENCH_ARGS="--spp 64" tools/variants.sh run exact ds fm base
