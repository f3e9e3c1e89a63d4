#!/usr/bin/env python
"""The BASELINE.json correctness gates at FULL size (run on the GPU box, report kept
under profiles/): the hit-id gate on 2^22 dumped primary rays + the shadow rays of the
same pixels at samples {0,1,17}, the image gate at the given spp (tests/gates.py; the
`gpu` test suite runs the same code at 1080p / 64 spp).
usage: python tools/full_gates.py <workload> <width> <height> <spp> [out.json]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gates  # noqa: E402
from raytracingproject_b200 import scenes  # noqa: E402


def main():
    workload, w, h, spp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    out_path = sys.argv[5] if len(sys.argv) > 5 else "gpurun_out/gates_%s.json" % workload
    make = {"terrain": scenes.terrain, "instanced": scenes.instanced, "cornell": scenes.cornell,
            "cube": scenes.default_cube}[workload]
    report = gates.run(make(width=w, height=h), spp)
    for rec in report["hit_id_gate"]:
        print(rec, flush=True)
    print(json.dumps(report["image_gate"]), "PASS" if report["pass"] else "FAIL", flush=True)
    with open(out_path, "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
