#!/usr/bin/env python3
"""Static SASS mnemonic histogram of one kernel of libpanacus_b200.so:  tools/sass_fn.py <substring of the mangled name>"""
import collections, re, subprocess, sys
so = "panacus_b200/libpanacus_b200.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, hist = None, collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and sys.argv[1] in cur:
        m = re.search(r"/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            hist[m.group(2).split(".")[0] + ("." + ".".join(m.group(2).split(".")[1:3]) if "." in m.group(2) else "")] += 1
for k, v in hist.most_common(40):
    print(f"{v:6d} {k}")
print("total", sum(hist.values()))
