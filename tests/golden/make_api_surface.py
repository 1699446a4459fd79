#!/usr/bin/env python3
"""Extracts the public surface of the reference's hot-path types into tests/golden/reference_api_surface.json (names only).
Run in the development container (needs /root/reference); the test that uses the JSON does not.

    python tests/golden/make_api_surface.py
"""
import json
import os
import re

REF = "/root/reference/SdfKit"
HERE = os.path.dirname(os.path.abspath(__file__))
# reference type -> (file, why it is in scope); SURVEY.md section 8a/8f
TYPES = {
    "SdfConfig": "Sdf.cs", "SdfEx": "Sdf.cs", "SdfExprs": "SdfExpr.cs", "SdfExprEx": "SdfExpr.cs", "SdfIndexedInput": "SdfExpr.cs",
    "Voxels": "Voxels.cs", "MarchingCubes": "MarchingCubes.cs", "Mesh": "Mesh.cs", "RayMarcher": "RayMarcher.cs",
}


def members(path, type_name):
    text = open(os.path.join(REF, path)).read()
    m = re.search(r"public\s+(?:static\s+)?(?:class|struct)\s+%s\b[^{]*\{" % type_name, text)
    depth, i, start = 1, m.end(), m.end()
    while depth:
        depth += {"{": 1, "}": -1}.get(text[i], 0)
        i += 1
    body = text[start:i]
    names = set()
    depth = 0
    for line in body.splitlines():
        if depth == 0:
            mm = re.match(r"\s*public\s+(?:static\s+|readonly\s+|const\s+|override\s+)*[\w<>\[\],.?]+\s+(\w+)\s*(?:\(|\{|=|;|=>)", line)
            if mm and mm.group(1) not in ("this", type_name):
                names.add(mm.group(1))
            elif re.match(r"\s*public\s+[\w<>\[\],.?]+\s+this\s*\[", line):
                names.add("this[]")
        depth += line.count("{") - line.count("}")
    return sorted(names)


if __name__ == "__main__":
    out = {t: {"file": "SdfKit/" + f, "members": members(f, t)} for t, f in TYPES.items()}
    with open(os.path.join(HERE, "reference_api_surface.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")
    for t, v in out.items():
        print(t, v["members"])
