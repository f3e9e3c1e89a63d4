import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["cornell", "terrain", "instanced", "cube"]


def load_traverse(name):
    z = np.load(os.path.join(GOLDEN, "traverse_%s.npz" % name))
    arrays = {k: z[k] for k in z.files if k.startswith("__")}
    return arrays, z["rays"], z["hits"]


def load_film(name):
    z = np.load(os.path.join(GOLDEN, "film_%s.npz" % name))
    return z["film"], int(z["spp"]), z["ray_counts"]


def golden_descs():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden",
                                                  os.path.join(GOLDEN, "make_golden.py"))
    # only the case table is needed; importing make_golden would import the oracle
    from raytracingproject_b200 import scenes
    W, H = 48, 27
    return {
        "cornell": scenes.cornell(W, H, materials="diffuse"),
        "terrain": scenes.terrain(W, H, n=24),
        "instanced": scenes.instanced(W, H, grid=4, subdiv=2),
        "cube": scenes.default_cube(W, H, material="diffuse"),
    }
