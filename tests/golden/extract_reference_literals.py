#!/usr/bin/env python
"""Extracts the literals of the reference's ray-tracing sample that define the path's INPUTS (scene, camera, shader constants) from
vulkan-raytracing-basic/main.cpp into tests/golden/reference_literals.json. These are the only golden values the reference holds for
this path (it ships no expected outputs). Run in the build container, where /root/reference exists:
    python tests/golden/extract_reference_literals.py [/root/reference]
The committed JSON travels to the GPU box; tests/test_reference_literals.py compares our scene description with it (and re-extracts when
the reference tree is present)."""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def _floats(text):
    return [float(x.rstrip("f")) for x in re.findall(r"-?\d+\.\d*f?|-?\d+f?", text)]


def extract(ref_root="/root/reference"):
    path = os.path.join(ref_root, "vulkan-raytracing-basic", "main.cpp")
    src = open(path).read()
    twin = open(os.path.join(ref_root, "vulkan-raytraced-triangle", "main.cpp")).read()
    out = {"source": "vulkan-raytracing-basic/main.cpp", "identical_to_vulkan_raytraced_triangle": src == twin}
    out["width"] = int(re.search(r"const uint32_t WIDTH = (\d+);", src).group(1))
    out["height"] = int(re.search(r"const uint32_t HEIGHT = (\d+);", src).group(1))
    m = re.search(r"float vertices\[\]\[3\] = \{(.*?)\};\s*uint32_t indices\[\] = \{(.*?)\};", src, re.S)
    v = _floats(m.group(1))
    out["vertices"] = [v[i:i + 3] for i in range(0, len(v), 3)]
    out["indices"] = [int(x) for x in re.findall(r"\d+", m.group(2))]
    g = re.search(r"VkTransformMatrixKHR geoTransforms\[\] = \{(.*?)\};", src, re.S)
    gv = _floats(g.group(1))
    out["geometry_transforms"] = [gv[i:i + 12] for i in range(0, len(gv), 12)]
    t = re.search(r"VkTransformMatrixKHR insTransforms\[\] = \{(.*?)\};", src, re.S)
    tv = _floats(t.group(1))
    out["instance_transforms"] = [tv[i:i + 12] for i in range(0, len(tv), 12)]
    out["instance_custom_index"] = int(re.search(r"\.instanceCustomIndex = (\d+)", src).group(1))
    out["instance_mask"] = int(re.search(r"\.mask = (0x[0-9A-Fa-f]+)", src).group(1), 16)
    out["instance_sbt_offsets"] = [int(re.search(r"\.instanceShaderBindingTableRecordOffset = (\d+)", src).group(1)),
                                   int(re.search(r"instanceData\[1\]\.instanceShaderBindingTableRecordOffset = (\d+)", src).group(1))]
    out["instance_flags"] = re.search(r"\.flags = (VK_GEOMETRY_INSTANCE_[A-Z_]+)", src).group(1)
    cam = _floats(re.search(r"\*\(Data\*\) dst = \{(.*?)\};", src).group(1))
    out["camera_pos"], out["yfov_deg"] = cam[:3], cam[3]
    out["hit_records"] = [_floats(x) for x in re.findall(r"\(HitgCustomData\*\s*\)\(dst \+ hitgOffset \+ \d \* hitgStride \+ handleSize\) = \{(.*?)\};", src)]
    out["miss_color"] = _floats(re.search(r"hitValue = vec3\((0\.0, 0\.0, 0\.2)\);", src).group(1))
    tr = re.search(r"gl_RayFlagsOpaqueEXT, (0x[0-9a-f]+),.*?\n.*?\n\s*g\.cameraPos, ([\d.]+), rayDir, ([\d.]+),", src)
    out["cull_mask"], out["tmin"], out["tmax"] = int(tr.group(1), 16), float(tr.group(2)), float(tr.group(3))
    sp = re.search(r"if \(gl_PrimitiveID == (\d+) &&\s*gl_InstanceID == (\d+) &&\s*gl_InstanceCustomIndexEXT == (\d+) &&\s*gl_GeometryIndexEXT == (\d+)\)", src)
    out["barycentric_case"] = {"primitive": int(sp.group(1)), "instance": int(sp.group(2)), "custom_index": int(sp.group(3)), "geometry": int(sp.group(4))}
    return out


if __name__ == "__main__":
    lit = extract(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    with open(os.path.join(HERE, "reference_literals.json"), "w") as f:
        json.dump(lit, f, indent=1)
    print(json.dumps(lit, indent=1))
