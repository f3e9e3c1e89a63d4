#!/usr/bin/env python
"""Generates the committed golden vectors from the REFERENCE itself
(oracle/_ref/libcycles_ref.so = the reference CPU Cycles compiled from
/root/reference).  The reference ships no known-answer tests for this path
(SURVEY.md 4), so the pins are outputs of the reference run here:

  traverse_<scene>.npz  packed BVH2 arrays as handed to a Device + dumped camera and
                        shadow rays + the reference's scene_intersect results
  film_<scene>.npz      the reference's film (generic scalar kernel) for a tiny render

usage (needs /root/reference + `make -C oracle`):  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import cycles_ref  # noqa: E402
from raytracingproject_b200 import scenes  # noqa: E402

BVH_ARRAYS = ["__bvh_nodes", "__bvh_leaf_nodes", "__prim_tri_verts", "__prim_tri_index",
              "__prim_type", "__prim_visibility", "__prim_index", "__prim_object",
              "__object_node", "__objects", "__object_flag", "__tri_shader", "__tri_vindex"]
W, H, SPP = 48, 27, 4


def golden_cases():
    return {
        "cornell": scenes.cornell(W, H, materials="diffuse"),
        "terrain": scenes.terrain(W, H, n=24),
        "instanced": scenes.instanced(W, H, grid=4, subdiv=2),
        "cube": scenes.default_cube(W, H, material="diffuse"),
    }


def main():
    for name, desc in golden_cases().items():
        rs = cycles_ref.build_scene(desc, kernel=cycles_ref.RefScene.GENERIC)
        arrays = rs.device_arrays()
        out = {"__data": arrays["__data"][0]}
        for a in BVH_ARRAYS:
            if a in arrays:
                out[a] = arrays[a][0]
        rays, hits = [], []
        for sample in (0, 1, 5):
            r, _ = rs.camera_rays(sample, 0, 0, W, H)
            rays.append(r)
            hits.append(rs.intersect(r))
            s = rs.shadow_rays(sample, 0, 0, W, H)
            rays.append(s)
            hits.append(rs.intersect(s))
        out["rays"] = np.concatenate(rays)
        out["hits"] = np.concatenate(hits)
        np.savez_compressed(os.path.join(HERE, "traverse_%s.npz" % name), **out)
        film, _ = rs.render(0, SPP, tile_size=16)
        counts = np.array(rs.count_rays(0, SPP), dtype=np.uint64)
        np.savez_compressed(os.path.join(HERE, "film_%s.npz" % name), film=film, spp=SPP,
                            ray_counts=counts)
        print(name, "rays", len(out["rays"]), "hit", int((out["hits"]["prim"] >= 0).sum()),
              "film mean", float(film[..., :3].mean() / SPP), "rays", counts)
        rs.close()


if __name__ == "__main__":
    main()
