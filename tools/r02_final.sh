#!/bin/bash
# end-of-round evidence on one B200 box: GPU tests, both bench arms, ncu launch list and
# --set full captures of the dominant kernels (same build)
cd "$(dirname "$0")/.."
O=gpurun_out
T=${1:-r02k}
( time python -m pytest tests -m gpu -q ) > $O/${T}_pytest.log 2>&1
grep -E "passed|failed" $O/${T}_pytest.log | tail -2
python __graft_entry__.py smoke > $O/${T}_smoke.log 2>&1; tail -2 $O/${T}_smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_reference_arm.json 2> $O/${T}_ref.err
python bench.py --steps 3 --warmup 3 > $O/${T}_bench_n1.json 2> $O/${T}_bench.err
tail -c 300 $O/${T}_bench.err
B="--no-cpu-baseline --no-e2e --configs none"
python bench.py --steps 2 --warmup 1 $B > $O/${T}_b_plain.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${T}_launches.csv \
  python bench.py --steps 2 --warmup 1 $B > $O/${T}_ncu_ll.log 2>&1
# the reports stay on the box (gpurun_out is capped at 64 MiB): summaries come back
R=/tmp/ncu_reports; mkdir -p $R
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_intersect_closest --launch-count 2 \
  -o $R/${T}_closest_terrain -f python bench.py --spp 16 --steps 1 --warmup 0 $B > $O/${T}_ncu_a1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_intersect_closest --launch-count 3 \
  -o $R/${T}_closest_instanced -f python bench.py --workload instanced --spp 4 --steps 1 --warmup 0 $B > $O/${T}_ncu_a2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shade_surface --launch-count 2 \
  -o $R/${T}_shade_cornell -f python bench.py --workload cornell --spp 16 --steps 1 --warmup 0 $B > $O/${T}_ncu_a3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shade_surface --launch-count 2 \
  -o $R/${T}_shade_cube_dense -f python bench.py --workload cube --spp 16 --steps 1 --warmup 0 --opt shade_wide=1 $B > $O/${T}_ncu_a4.log 2>&1
for n in closest_terrain closest_instanced shade_cornell shade_cube_dense; do
  python tools/ncu_summary.py $R/${T}_$n.ncu-rep > $O/${T}_prof_$n.txt 2>&1
  python tools/srclines.py $R/${T}_$n.ncu-rep 0 40 > $O/${T}_srclines_$n.txt 2>&1
done
python tools/ncu_traffic.py $T terrain=$R/${T}_closest_terrain.ncu-rep instanced=$R/${T}_closest_instanced.ncu-rep > $O/${T}_traffic.log 2>&1
cp profiles/traffic_ncu.json $O/${T}_traffic_ncu.json
ls -la $O | grep ${T}; du -sh $O
