"""Histogram of SASS opcodes per kernel of an object file / library: python tools/sass_ops.py <file> [name filter]."""
import collections
import re
import subprocess
import sys

txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
flt = sys.argv[2] if len(sys.argv) > 2 else ""
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n")[0]
    if flt not in name:
        continue
    ops = collections.Counter()
    for line in f.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops[m.group(2).split(".")[0]] += 1
    print(name[:150])
    print("  total", sum(ops.values()), dict(ops.most_common(30)))
