"""The BASELINE.json correctness gates at configuration size, as a function the `gpu`
tests (tests/test_full_gates_gpu.py) and tools/full_gates.py share:

  - hit-id gate: bit-exact prim / object ids on dumped primary rays and the shadow rays
    of the same pixels against the reference's scene_intersect, |dt| < 1e-5 * t excluded
    AND counted;
  - image gate: per-pixel RMSE <= 1e-3 and mean luminance within 0.1 % against the
    reference generic (parity) CPU kernel on the same seeds and samples.

TEST INFRASTRUCTURE: uses oracle/ (the reference compiled here) as the checker."""
import time

import numpy as np


def exact_hit(arrays, ray, hit):
    """Double-precision ray/triangle test of the triangle a hit record names (in object
    space for instances).  Used only to CLASSIFY the rare hard mismatches: the
    reference's float test accepts some near-parallel triangles (U = V = 0 by
    cancellation, |den| ~ 1e-7) that the ray geometrically misses."""
    prim, obj = int(hit["prim"]), int(hit["object"])
    if prim < 0:
        return False
    tv = arrays["__prim_tri_verts"][0].view(np.float32).reshape(-1, 4)
    ti = arrays["__prim_tri_index"][0].view(np.uint32)
    a, b, c = tv[ti[prim]:ti[prim] + 3, :3].astype(np.float64)
    P, D = ray["P"].astype(np.float64), ray["D"].astype(np.float64)
    if obj >= 0:
        itfm = arrays["__objects"][0].view(np.float32).reshape(-1, 48)[obj][12:24].reshape(3, 4)
        itfm = itfm.astype(np.float64)
        P, D = itfm[:, :3] @ P + itfm[:, 3], itfm[:, :3] @ D
    n = np.cross(b - a, c - a)
    den = D @ n
    if den == 0:
        return False
    t = ((a - P) @ n) / den
    hp = P + D * t
    v0, v1, v2 = b - a, c - a, hp - a
    d00, d01, d11, d20, d21 = v0 @ v0, v0 @ v1, v1 @ v1, v2 @ v0, v2 @ v1
    dd = d00 * d11 - d01 * d01
    u, v = (d11 * d20 - d01 * d21) / dd, (d00 * d21 - d01 * d20) / dd
    tol = 1e-6
    return bool(t > 0 and u >= -tol and v >= -tol and u + v <= 1 + tol)


def hit_id_gate(rs, dev, arrays, w, h, samples=(0, 1, 17), max_rays=1 << 22):
    """One record per sample index; `hard_mismatches` must be 0."""
    out = []
    rows = min(h, max(1, max_rays // w))
    for sample in samples:
        rays, _ = rs.camera_rays(sample, 0, 0, w, rows)
        ref_hits = rs.intersect(rays)
        got = dev.trace_batch(rays)
        same = (ref_hits["prim"] == got["prim"]) & (ref_hits["object"] == got["object"])
        rel = np.abs(ref_hits["t"] - got["t"]) / np.maximum(np.abs(ref_hits["t"]), 1e-30)
        grazing = ~same & (ref_hits["prim"] >= 0) & (got["prim"] >= 0) & (rel < 1e-5)
        hard = np.nonzero(~same & ~grazing)[0]
        ref_false_pos = 0
        for i in hard:
            if not exact_hit(arrays, rays[i], ref_hits[i]) and (
                    got["prim"][i] < 0 or exact_hit(arrays, rays[i], got[i])):
                ref_false_pos += 1
        srays = rs.shadow_rays(sample, 0, 0, w, rows)
        ref_occ = rs.intersect(srays)["prim"] >= 0
        got_occ = dev.trace_batch(srays, any_hit=True)["prim"] >= 0
        m = same & (ref_hits["prim"] >= 0)
        out.append({
            "sample": sample, "primary_rays": int(len(rays)),
            "primary_hits": int((ref_hits["prim"] >= 0).sum()),
            "id_mismatches": int((~same).sum()), "grazing_excluded": int(grazing.sum()),
            "hard_mismatches": int(len(hard)) - ref_false_pos,
            "reference_false_positives": ref_false_pos,
            "uv_bit_identical": bool(np.array_equal(ref_hits["u"][m], got["u"][m]) and
                                     np.array_equal(ref_hits["v"][m], got["v"][m])),
            "shadow_rays": int((srays["t"] != 0).sum()), "shadow_occluded": int(ref_occ.sum()),
            "shadow_mismatches": int((ref_occ != got_occ).sum())})
    return out


def luminance(im):
    return 0.2126 * im[..., 0] + 0.7152 * im[..., 1] + 0.0722 * im[..., 2]


def image_gate(rs, dev, w, h, spp):
    ref_img, cpu_s = rs.render(0, spp, tile_size=64)
    got = dev.render(w, h, rs.pass_stride, 0, spp)
    st = dev.stats()
    a = ref_img[..., :3].astype(np.float64) / spp
    b = got[..., :3].astype(np.float64) / spp
    la, lb = luminance(a).mean(), luminance(b).mean()
    return {
        "rmse": float(np.sqrt(np.mean((a - b) ** 2))),
        "mean_luminance_ref": float(la), "mean_luminance_b200": float(lb),
        "mean_luminance_rel_diff": float(abs(la - lb) / la),
        "max_abs_diff": float(np.abs(a - b).max()),
        "bit_identical_pixels": float(np.mean(np.all(ref_img == got, axis=-1))),
        "alpha_identical": bool(np.array_equal(ref_img[..., 3], got[..., 3])),
        "cpu_generic_kernel_s": cpu_s, "gpu_device_ms": st["device_ms"],
        "gpu_rays": {k: st[k] for k in ("primary_rays", "bounce_rays", "shadow_rays")}}


def run(desc, spp, samples=(0, 1, 17), max_rays=1 << 22, device=None):
    """Both gates on one scene description; returns the report with `pass`."""
    from oracle import cycles_ref as cr
    from raytracingproject_b200.device import B200Device
    t0 = time.time()
    rs = cr.build_scene(desc, kernel=cr.RefScene.GENERIC)
    dev = device or B200Device(0)
    try:
        arrays = rs.device_arrays()
        dev.upload_scene(arrays)
        w, h = desc.width, desc.height
        report = {"workload": desc.name, "width": w, "height": h, "spp": spp,
                  "triangles": desc.num_triangles, "objects": len(desc.objects),
                  "bvh8": dev.build_bvh(), "scene_s": time.time() - t0}
        report["hit_id_gate"] = hit_id_gate(rs, dev, arrays, w, h, samples, max_rays)
        report["image_gate"] = image_gate(rs, dev, w, h, spp)
        g = report["image_gate"]
        report["pass"] = bool(g["rmse"] <= 1e-3 and g["mean_luminance_rel_diff"] <= 1e-3 and
                              all(x["hard_mismatches"] == 0 and x["shadow_mismatches"] == 0
                                  for x in report["hit_id_gate"]))
        return report
    finally:
        if device is None:
            dev.close()
        rs.close()
