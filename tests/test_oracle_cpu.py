"""CPU suite (-m "not gpu"): the oracles against the golden vectors dumped from the
reference (tests/golden/make_golden.py), and the product's host BVH8 builder."""
import numpy as np
import pytest

import golden_util as G
import hostcheck


def assert_same_hits(a, b):
    for f in ("prim", "object", "type"):
        assert np.array_equal(a[f], b[f]), f
    for f in ("t", "u", "v"):
        assert np.array_equal(a[f].view(np.uint32), b[f].view(np.uint32)), f


@pytest.mark.parametrize("name", G.CASES)
def test_c_restatement_matches_reference_golden(name):
    """oracle/cycles_port.c (plain C) == the reference's scene_intersect, bit for bit."""
    from oracle.cycles_port import PortOracle
    arrays, rays, hits = G.load_traverse(name)
    got = PortOracle(arrays).intersect(rays)
    assert_same_hits(hits, got)


@pytest.mark.parametrize("name", G.CASES)
def test_reference_build_reproduces_golden(ref, name):
    """oracle/_ref (the reference compiled here) still produces the committed vectors:
    same packed BVH2 arrays, same hits, same film, same ray census."""
    desc = G.golden_descs()[name]
    rs = ref.build_scene(desc, kernel=ref.RefScene.GENERIC)
    try:
        arrays, rays, hits = G.load_traverse(name)
        live = rs.device_arrays()
        for k, v in arrays.items():
            if k == "__data":
                continue  # carries pointers / padding that differ between runs
            a, b = live[k][0], v
            if k == "__objects":  # KernelObject ends in uninitialised padding: tfm + itfm only
                a, b = a.reshape(-1, 192)[:, :96], b.reshape(-1, 192)[:, :96]
            assert np.array_equal(a, b), k
        assert_same_hits(hits, rs.intersect(rays))
        film, spp, counts = G.load_film(name)
        got, _ = rs.render(0, spp, tile_size=16)
        assert np.array_equal(got, film)
        assert tuple(int(c) for c in counts) == rs.count_rays(0, spp)
    finally:
        rs.close()


@pytest.mark.parametrize("name", G.CASES)
def test_host_bvh8_builder_against_golden(name):
    """The product's host BVH8 builder (csrc/bvh8_build.cpp), walked on the CPU by the
    test-only checker: structure invariants + closest-hit ids equal to the reference."""
    arrays, rays, hits = G.load_traverse(name)
    hb = hostcheck.HostBVH8({k: (v, 1) for k, v in arrays.items()})
    info = hb.info()
    bad, counts = hb.invariants()
    assert bad == 0
    is_tri = arrays["__prim_index"].view(np.int32) != -1
    assert np.all(counts[is_tri] == 1) and np.all(counts[~is_tri] == 0)
    assert info["triangles"] == int(is_tri.sum())
    got = hb.intersect(rays)
    # shadow rays (PATH_RAY_SHADOW_OPAQUE) stop at the FIRST hit found, which depends on
    # traversal order: only occlusion is comparable for them
    shadow = (rays["visibility"] & 0x180) != 0
    assert np.array_equal(got["prim"][shadow] >= 0, hits["prim"][shadow] >= 0)
    got, hits = got[~shadow], hits[~shadow]
    same = (got["prim"] == hits["prim"]) & (got["object"] == hits["object"])
    rel = np.abs(got["t"] - hits["t"]) / np.maximum(np.abs(hits["t"]), 1e-30)
    grazing = ~same & (got["prim"] >= 0) & (hits["prim"] >= 0) & (rel < 1e-5)
    assert (~same & ~grazing).sum() == 0
    hit = same & (hits["prim"] >= 0)
    assert np.array_equal(got["u"][hit], hits["u"][hit])
    assert np.array_equal(got["v"][hit], hits["v"][hit])


def test_host_bvh8_rejects_out_of_scope():
    """Curves / motion leaves are refused by the builder, not mis-traversed."""
    arrays, _, _ = G.load_traverse("cube")
    arrays = dict(arrays)
    leaves = arrays["__bvh_leaf_nodes"].copy().view(np.uint32).reshape(-1, 4)
    leaves[:, 3] = 0x20  # not PRIMITIVE_TRIANGLE
    arrays["__bvh_leaf_nodes"] = leaves.view(np.uint8).reshape(-1)
    with pytest.raises(RuntimeError):
        hostcheck.HostBVH8({k: (v, 1) for k, v in arrays.items()})


def test_empty_ray_batches():
    from oracle.cycles_port import PortOracle
    arrays, rays, _ = G.load_traverse("cube")
    assert len(PortOracle(arrays).intersect(rays[:0])) == 0
    hb = hostcheck.HostBVH8({k: (v, 1) for k, v in arrays.items()})
    assert len(hb.intersect(rays[:0])) == 0
    dead = rays[:16].copy()
    dead["t"] = 0.0
    assert (PortOracle(arrays).intersect(dead)["prim"] == -1).all()
