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
