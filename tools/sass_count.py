#!/usr/bin/env python
"""Static SASS mnemonic histogram per kernel of an object file / shared library.
usage: tools/sass_count.py file.o [substring of the demangled kernel name ...]"""
import collections, re, subprocess, sys
obj, pats = sys.argv[1], sys.argv[2:]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        hist[cur][m.group(2).split(".")[0]] += 1
for k, h in hist.items():
    if pats and not all(p in k for p in pats):
        continue
    tot = sum(h.values())
    print(f"{k[:150]}\n  total {tot}: " + " ".join(f"{a}={b}" for a, b in h.most_common(28)))
