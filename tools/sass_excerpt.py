"""profiles/sass_tma_excerpt.txt: per kernel, the TMA (UBLKCP) and mbarrier (SYNCS) instructions in the SASS of the
built library, and the count of tensor-core instructions (none: the path is byte and integer work).  No GPU needed."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "agent0_b200", "libagent0_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0] or n
out = ["SASS evidence of the TMA path (cuobjdump -sass agent0_b200/libagent0_b200.so, sm_100a; regenerate with tools/sass_excerpt.py).",
       "UBLKCP = cp.async.bulk (1-D TMA bulk copy; .S.G global->shared, .G.S shared->global), SYNCS = mbarrier operations.",
       "No tensor-core instruction (UTC*MMA / HMMA) exists anywhere in the library: the path is byte and integer work.", ""]
cur, per = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = []
        continue
    if cur and re.search(r"UBLKCP|SYNCS|UTC\w*MMA|HMMA|UTMALDG", line):
        per[cur].append(line.strip())
for fn, ls in per.items():
    if not ls:
        continue
    c = collections.Counter(re.search(r"(UBLKCP\S*|SYNCS\S*|UTC\S*|HMMA\S*|UTMALDG\S*)", l).group(1) for l in ls)
    out.append(f"{demangle(fn)}: " + ", ".join(f"{k} x{v}" for k, v in c.items()))
    for l in ls[:4]:
        out.append("    " + re.sub(r"\s+/\*[0-9a-fx]+\*/\s*$", "", l))
    if len(ls) > 4:
        out.append(f"    ... {len(ls) - 4} more")
out += ["", f"kernels in the library: {len(per)}; with TMA/mbarrier instructions: {sum(1 for v in per.values() if v)}; "
            f"tensor-core instructions: {sum(1 for v in per.values() for l in v if re.search('MMA', l))}"]
open(os.path.join(ROOT, "profiles", "sass_tma_excerpt.txt"), "w").write("\n".join(out) + "\n")
sys.stdout.write("\n".join(out[-8:]) + "\n")
