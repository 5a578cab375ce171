"""The C-ABI library loads on a GPU-less host and exports exactly what include/*.h declare."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    inc = os.path.join(ROOT, "include")
    for h in sorted(os.listdir(inc)):
        if not h.endswith(".h"):
            continue
        src = open(os.path.join(inc, h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names.update(re.findall(r"RT_API[^;(]*?\b(rt_[a-z_0-9]+)\s*\(", src))
    return sorted(names)


def test_header_symbols_exported(rt):
    L = rt.load()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/*.h but not exported by librtcore.so"
    assert sorted(rt.EXPORTED_SYMBOLS) == names


def test_struct_layouts(rt):
    assert C.sizeof(rt.RtInstance) == 64          # VkAccelerationStructureInstanceKHR
    assert C.sizeof(rt.RtCamera) == 16            # std140 vec3 + float
    assert rt.HIT_DTYPE.itemsize == 28
    assert C.sizeof(rt.RtGeometry) == 48


def test_no_cpu_fallback(rt):
    """Without a CUDA device the product refuses to work instead of falling back to anything."""
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    with pytest.raises(rt.RtError):
        rt.Context(0)


def test_product_does_not_reference_oracle():
    pkg = os.path.join(ROOT, "build-up-phase_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_binding" not in text and "liboracle" not in text and "rt_oracle" not in text, f
