"""Generates tests/golden/reference_default_xml.json from the reference's own config/default.xml (the one artefact of
the reference that pins anything on this path: the default parameter values). Run in the build container, where
/root/reference exists; the fixture is committed so that the test needs no access to the reference."""
import json
import os
import sys
import xml.etree.ElementTree as ET

import re

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/config/default.xml"
HERE = os.path.dirname(os.path.abspath(__file__))
# only the keys the hot path reads: the ones suma::ParameterList maps onto sb_params, plus the weighting string
hpp = open(os.path.join(HERE, "..", "..", "include", "suma_b200.hpp")).read()
wanted = set(re.findall(r'SUMA_F\("([^"]+)"', hpp)) | {"weighting"}
out = {}
for prm in ET.parse(SRC).getroot().findall("param"):
    name, typ, txt = prm.get("name"), prm.get("type"), (prm.text or "").strip()
    if name not in wanted:
        continue
    if typ == "integer":
        val = int(txt)
    elif typ == "float":
        val = float(txt)
    elif typ == "boolean":
        val = txt.lower() == "true"
    else:
        val = txt
    out[name] = {"type": typ, "value": val}
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_default_xml.json")
json.dump({"source": "PRBonn/semantic_suma config/default.xml", "params": out}, open(dst, "w"), indent=1, sort_keys=True)
print("wrote", dst, len(out), "parameters")
