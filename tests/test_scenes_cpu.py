"""CPU suite: the procedural workloads are deterministic and have the stated sizes."""
import hashlib

import numpy as np

from raytracingproject_b200 import scenes


def digest(desc):
    h = hashlib.sha256(desc.xml.encode())
    for m in desc.meshes:
        h.update(np.ascontiguousarray(m.P).tobytes())
        h.update(np.ascontiguousarray(m.tris).tobytes())
    for mi, t in desc.objects:
        h.update(np.ascontiguousarray(t, dtype=np.float32).tobytes())
    return h.hexdigest()


def test_generators_are_deterministic():
    for make in (lambda: scenes.terrain(64, 36, n=32), lambda: scenes.cornell(64, 36),
                 lambda: scenes.instanced(64, 36, grid=3, subdiv=1),
                 lambda: scenes.default_cube(64, 36)):
        assert digest(make()) == digest(make())


def test_config_sizes():
    t = scenes.terrain(n=708)
    assert t.num_triangles == 1002528 and (t.width, t.height, t.spp) == (1920, 1080, 256)
    i = scenes.instanced(grid=5, subdiv=2)
    assert len(i.objects) == 26 and i.meshes[0].tris.shape == (320, 3)
    assert i.num_instanced_triangles == 25 * 320 + 2
    c = scenes.default_cube()
    assert c.num_triangles == 12 and (c.width, c.height, c.spp) == (1920, 1080, 64)
    full = scenes.instanced()
    assert (full.width, full.height) == (3840, 2160) and len(full.objects) == 10001
    assert full.meshes[0].tris.shape[0] == 100820  # 20 * 71^2: BASELINE config 4's 100k-triangle BLAS


def test_meshes_are_valid():
    for d in (scenes.terrain(64, 36, n=16), scenes.instanced(64, 36, grid=2, subdiv=2),
              scenes.cornell(64, 36)):
        for m in d.meshes:
            assert m.tris.min() >= 0 and m.tris.max() < len(m.P)
            assert np.isfinite(m.P).all()
