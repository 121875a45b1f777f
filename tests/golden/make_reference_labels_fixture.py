"""Generates tests/golden/reference_movable_labels.json from the reference's shaders: the label ids each pass treats
as "movable" (gen_vertexmap.vert:95-100 removes them during the first scans, update_surfels.vert:187-195 penalises
them, gen_surfels.geom:135-140 lowers their initial confidence). Run where /root/reference exists."""
import json
import os
import re
import sys

ROOT = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/shader"
def class_ids(src):
    return {m.group(1): float(m.group(2)) for m in
            re.finditer(r"vec4\s+(\w+)\s*=\s*vec4\(\s*[\d.]+\s*,\s*[\d.]+\s*,\s*[\d.]+\s*,\s*([\d.]+)\s*\)", src)}


shared = class_ids(open(os.path.join(ROOT, "color_map.glsl")).read())  # #include'd by update_surfels.vert
out = {}
for fn in ("gen_vertexmap.vert", "update_surfels.vert", "gen_surfels.geom"):
    src = open(os.path.join(ROOT, fn)).read()
    src = re.sub(r"//[^\n]*", "", src)
    ids = dict(shared)
    ids.update(class_ids(src))
    conds = [c for c in re.findall(r"if\s*\(([^{;]*?_label\s*==[^{;]*?)\)\s*[\n{a-zA-Z]", src, flags=re.S) if ".w" in c]
    movable = sorted({ids[n] for c in conds for n in re.findall(r"==\s*(\w+)\.w", c)})
    out[fn] = {"class_ids": ids, "movable": movable}
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_movable_labels.json")
json.dump({"source": "PRBonn/semantic_suma src/shader", "passes": out}, open(dst, "w"), indent=1, sort_keys=True)
for k, v in out.items():
    print(k, v["movable"])
