#!/bin/bash
# evidence for the round's last build (tag r02t): tests, smoke, both bench arms, and the
# wide shading kernel on the startup scene under ncu
cd "$(dirname "$0")/.."
O=gpurun_out; T=r02u
( time python -m pytest tests -m gpu -q ) > $O/${T}_pytest.log 2>&1
grep -E "passed|failed" $O/${T}_pytest.log | tail -2
python __graft_entry__.py smoke > $O/${T}_smoke.log 2>&1; tail -1 $O/${T}_smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_reference_arm.json 2> $O/${T}_ref.err
python bench.py --steps 3 --warmup 3 > $O/${T}_bench_n1.json 2> $O/${T}_bench.err; tail -c 200 $O/${T}_bench.err
B="--no-cpu-baseline --no-e2e --configs none"
R=/tmp/ncu_reports; mkdir -p $R
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shade_surface --launch-count 2 \
  -o $R/${T}_shade_cube_wide -f python bench.py --workload cube --spp 16 --steps 1 --warmup 0 --opt shade_wide=2 $B > $O/${T}_ncu_a6.log 2>&1
python tools/ncu_summary.py $R/${T}_shade_cube_wide.ncu-rep > $O/${T}_prof_shade_cube_wide.txt 2>&1
python tools/srclines.py $R/${T}_shade_cube_wide.ncu-rep 0 40 > $O/${T}_srclines_shade_cube_wide.txt 2>&1
ls -la $O | grep ${T}
