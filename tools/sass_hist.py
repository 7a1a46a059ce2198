"""Opcode histogram of one kernel's SASS (static counts): python tools/sass_hist.py <obj|so> <kernel-substring>"""
import collections
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, hist = None, collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and pat in cur:
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            hist[m.group(2)] += 1
print(sum(hist.values()), "instructions")
for k, v in hist.most_common(25):
    print("%6d %s" % (v, k))
