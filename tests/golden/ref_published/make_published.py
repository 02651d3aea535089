"""Extracts the two result matrices the reference publishes in docs/src/index.md (lines 93 and 96: `A = [...]`, rows separated
by `;`) into index_md_statistics.json.  Run in the build container, where /root/reference exists; the JSON travels."""
import json
import os
import re
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/docs/src/index.md"
out = {}
for line in open(src):
    m = re.match(r"\s*#?\s*A = \[(.*)\]\s*$", line)
    if not m:
        continue
    rows = [[float(x) for x in row.split()] for row in m.group(1).split(";")]
    assert len(rows) == 7 and len({len(r) for r in rows}) == 1
    d = int(rows[1][0])
    out[str(d)] = {"source": "docs/src/index.md", "nodes": [int(x) for x in rows[0]], "vertices": rows[3], "boundary_vertices": rows[4],
                   "walks": rows[5], "nn_per_walk": rows[6]}
assert sorted(out) == ["4", "5"]
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "index_md_statistics.json"), "w") as f:
    json.dump(out, f)
print({k: len(v["nodes"]) for k, v in out.items()})
