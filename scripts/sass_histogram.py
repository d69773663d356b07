"""SASS evidence for the Blackwell-native kernels: disassembles the in-tree objects (adv_grpo_b200/csrc/build/*.o, built by
__graft_entry__.build()) with `cuobjdump -sass` and writes, per kernel, the counts of the instructions that prove the
tcgen05 / TMEM / TMA path (UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG /
UTMAREDG = cp.async.bulk.tensor load / store / reduce, UBLKCP = cp.async.bulk, SYNCS = mbarrier) to
profiles/<tag>_sass_histogram.txt.  No GPU needed.   Usage: python scripts/sass_histogram.py [tag]"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "SYNCS", "MUFU", "HMMA", "BAR"]
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
out = [f"# cuobjdump -sass of adv_grpo_b200/csrc/build/*.o (sm_100a), instruction counts per kernel ({', '.join(KEYS)})",
       "# .2CTA = cta_group::2 forms; HMMA = legacy mma.sync (must be 0 in the tensor-core kernels)", ""]
for obj in sorted(glob.glob(os.path.join(ROOT, "adv_grpo_b200", "csrc", "build", "*.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    fn, counts, details = None, {}, {}
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(anonymous namespace\)::|advgrpo::", "", fn)
            fn = re.sub(r"\(.*", "", fn)
            counts[fn] = collections.Counter()
            details[fn] = collections.Counter()
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]+)", line)
        if m and fn:
            op = m.group(1)
            base = op.split(".")[0]
            if base in KEYS:
                counts[fn][base] += 1
                if base in ("UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM"):
                    details[fn][op] += 1
    out.append(f"== {os.path.basename(obj)}")
    for fn in sorted(counts):
        c = counts[fn]
        if not any(c[k] for k in KEYS if k not in ("BAR", "MUFU", "SYNCS")):
            continue
        out.append(f"  {fn[:110]}")
        out.append("      " + "  ".join(f"{k}={c[k]}" for k in KEYS if c[k] or k == "HMMA"))
        out.append("      forms: " + ", ".join(f"{k} x{v}" for k, v in sorted(details[fn].items())))
    out.append("")
path = os.path.join(ROOT, "profiles", f"{tag}_sass_histogram.txt")
open(path, "w").write("\n".join(out) + "\n")
print(path)
