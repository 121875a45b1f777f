"""Generates tests/golden/reference_surfel_layout.json from the reference's core/Surfel.h: the member order and types of
the 64-byte record that getAllSurfels() hands to callers. Run where /root/reference exists."""
import json
import os
import re
import sys

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/core/Surfel.h"
body = re.search(r"struct\s+Surfel\s*\{(.*?)\};", open(SRC).read(), flags=re.S).group(1)
fields = []
for typ, names in re.findall(r"\b(float|uint32_t|int32_t)\s+([^;]+);", body):
    for n in names.split(","):
        fields.append([n.strip(), typ])
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_surfel_layout.json")
json.dump({"source": "PRBonn/semantic_suma src/core/Surfel.h", "fields": fields}, open(dst, "w"), indent=1)
print(fields)
