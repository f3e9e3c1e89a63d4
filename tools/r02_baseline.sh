#!/bin/bash
# round-2 "before" numbers of the round-1 build on this round's boxes
cd "$(dirname "$0")/.."
O=gpurun_out
B="--no-cpu-baseline --no-e2e"
python bench.py --steps 3 --warmup 3 $B > $O/r02base_terrain.json 2> $O/r02base_terrain.err
python bench.py --workload cornell --materials principled --steps 2 --warmup 1 $B > $O/r02base_cornell.json 2> $O/r02base_cornell.err
python bench.py --workload instanced --width 3840 --height 2160 --steps 2 --warmup 1 $B > $O/r02base_instanced.json 2> $O/r02base_instanced.err
python bench.py --workload cube --steps 2 --warmup 1 $B > $O/r02base_cube.json 2> $O/r02base_cube.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade_surface --launch-skip 3 --launch-count 2 \
  -o $O/r02base_shade -f python bench.py --workload cornell --materials principled --spp 32 --steps 1 --warmup 0 $B > $O/ncu_shade.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_intersect_closest --launch-skip 1 --launch-count 2 \
  -o $O/r02base_closest_inst -f python bench.py --workload instanced --width 3840 --height 2160 --spp 4 --steps 1 --warmup 0 $B > $O/ncu_closest_inst.log 2>&1
ls -la $O | tail -20
