"""CPU suite: the C-ABI library loads and exports every symbol include/b200_cycles.h
declares (no compute without a GPU), and fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "b200_cycles.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from raytracingproject_b200 import device
    lib = device.load_library()
    names = declared_functions()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(device.EXPORTS) == names
    assert lib.b200_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    import torch
    from raytracingproject_b200 import device
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(device.DeviceError) as e:
        device.B200Device(0)
    assert "CUDA" in str(e.value) or "device" in str(e.value)


def test_abi_offsets_header_is_consistent():
    text = open(os.path.join(ROOT, "include", "cycles_abi.h")).read()
    vals = dict(re.findall(r"#define (\w+)[ \t]+(\S+)", text))
    assert int(vals["SIZEOF_KERNEL_DATA"]) == 1584
    assert int(vals["SIZEOF_KERNEL_OBJECT"]) == 192
    assert int(vals["SIZEOF_KERNEL_LIGHT"]) == 192
    assert int(vals["KD_BVH_ROOT"]) % 4 == 0
    # golden KernelData blobs have the size the generated header says
    import numpy as np
    import golden_util as G
    for name in G.CASES:
        arrays, _, _ = G.load_traverse(name)
        assert arrays["__data"].size == int(vals["SIZEOF_KERNEL_DATA"])


def test_sm100_only_binary():
    """The shipped library carries sm_100a SASS and nothing else."""
    lib = os.path.join(ROOT, "raytracingproject_b200", "libb200cycles.so")
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("no cuobjdump")
    import subprocess
    out = subprocess.run([cuobjdump, "-lelf", lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_product_libraries_have_no_path_to_the_oracle_kernels():
    """The C-ABI library links only the CUDA runtime; the C++ shim links it and the
    reference HOST library (Device base, Scene ... - no kernels, no CPU device).  Neither
    depends on the oracle (libcycles_ref.so) nor references a kernel_cpu_* symbol."""
    import subprocess
    pkg = os.path.join(ROOT, "raytracingproject_b200")
    libs = [os.path.join(pkg, "libb200cycles.so")]
    shim = os.path.join(pkg, "libcycles_device_b200.so")
    host = os.path.join(ROOT, "oracle", "_ref", "libcycles_host.so")
    if os.path.exists(shim):
        libs.append(shim)
    if os.path.exists(host):
        libs.append(host)
    for lib in libs:
        needed = subprocess.run(["readelf", "-d", lib], capture_output=True, text=True).stdout
        assert "libcycles_ref" not in needed, lib
        syms = subprocess.run(["nm", "-D", lib], capture_output=True, text=True).stdout
        assert "kernel_cpu_" not in syms, lib
    needed = subprocess.run(["readelf", "-d", libs[0]], capture_output=True, text=True).stdout
    assert "libcycles" not in needed and "libnccl" not in needed  # NCCL is bound at first use


def test_bvh8_is_a_host_layout_of_the_reference(ref):
    """BVH_LAYOUT_BVH8 as a first-class host layout: the reference's own BVH::create +
    BVH::build, asked for that layout, reach `BVH8 : BVH` (csrc/bvh8_host.cpp, registered
    by the device library) and come back with the device's node / record arrays - byte for
    byte what b200_bvh8_pack makes of the packed binary tree, which is the BVH the device
    derives itself on a host that only knows BVH2 (and whose traversal the GPU parity tests
    check).  Host-only: no GPU involved."""
    import ctypes as C
    import numpy as np
    from raytracingproject_b200 import device as D, scenes
    D.load_library()
    if not os.path.exists(D.SHIM_PATH):
        pytest.skip("device shim not built (needs the reference headers)")
    C.CDLL(D.SHIM_PATH, mode=C.RTLD_GLOBAL)   # registers the layout with the host library
    for desc in (scenes.cornell(64, 36, materials="diffuse"),
                 scenes.instanced(64, 36, grid=4, subdiv=2),
                 scenes.terrain(64, 36, n=48)):
        rs = ref.build_scene(desc)
        try:
            arrays = rs.device_arrays()
            nodes, recs, onode, root, info = D.pack_bvh8(arrays)
            assert info["host_packed"] == 1 and info["num_nodes"] * 80 == nodes.size
            h_nodes, h_recs, h_onode, h_root = rs.pack_bvh(D.BVH_LAYOUT_BVH8)
            assert h_root == root
            assert np.array_equal(h_nodes, nodes) and np.array_equal(h_recs, recs)
            assert np.array_equal(h_onode[:len(onode)], onode)
            # and the layout the scene was built with is still the binary one
            b_nodes, _, _, b_root = rs.pack_bvh(D.BVH_LAYOUT_BVH2)
            assert np.array_equal(b_nodes, np.ascontiguousarray(arrays["__bvh_nodes"][0])) \
                if "__bvh_nodes" in arrays else b_nodes.size == 0
        finally:
            rs.close()


def test_bvh8_pack_refuses_what_the_device_cannot_traverse():
    import numpy as np
    from raytracingproject_b200 import device as D
    with pytest.raises(D.DeviceError, match="BVH2"):
        D.pack_bvh8({"__data": (np.zeros(4096, np.uint8), 1)})


def test_reference_registry_knows_the_device_type():
    """The registration patch applied to the host library's copy
    of device/device.cpp: "B200" is a DeviceType of the reference's registry, named both
    ways; without a GPU the plug-in reports no devices and Device::create gives NULL
    instead of another backend."""
    import os
    from raytracingproject_b200 import device as D
    if not os.path.exists(D.SHIM_PATH):
        pytest.skip("libcycles_device_b200.so not built (needs /root/reference)")
    reg = D.RegisteredDevice("B200", count=0)
    assert reg.type == 7 and reg.type_name() == "B200"       # DEVICE_OPTIX + 1
    assert D.RegisteredDevice("CPU", count=0).type == 1
    with pytest.raises(D.DeviceError):
        D.RegisteredDevice("B300", count=0)
    import torch
    if not torch.cuda.is_available():
        assert not reg.type_available() and reg.available() == []
        with pytest.raises(D.DeviceError):
            D.RegisteredDevice("B200", index=0)
