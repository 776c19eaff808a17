"""SASS evidence for profiles/: per kernel of libjz_b200.so, how many tcgen05 MMAs (UTC*MMA), TMEM loads (LDTM), TMA
loads (UTMALDG), bulk / async copies (UBLKCP, LDGSTS), system-scope stores (ST.SYS = STG.E.STRONG.SYS: what
multimem.st.relaxed.sys to an NVSwitch multicast address and the st.release.sys of the cross-GPU barrier compile to --
SASS has no separate mnemonic for a multicast store, the address decides), 128-bit global accesses (LDG.E.128 / STG.E.128), packed two-lane fp32 instructions (F*2 = FFMA2 / FADD2 / FMUL2), stores into another
CTA's shared memory (ST.E = generic stores: the st.shared::cluster of the cluster split-K exchange compile to these -- a mapa'd shared-window
address goes through the generic path -- as do the epilogue's stores through peer pointers), cluster barriers (UCGABAR = barrier.cluster arrive / wait) and legacy tensor instructions (HMMA: must be zero) appear.
    python scripts/sass_summary.py > profiles/sass_summary.txt      (build container: cuobjdump, no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "juzhen_b200", "libjz_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = collections.OrderedDict([
    ("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("UTMALDG", r"\bUTMALDG"), ("UBLKCP", r"\bUBLKCP"),
    ("LDGSTS", r"\bLDGSTS"), ("ST.SYS", r"\bSTG\.E\.STRONG\.SYS"), ("LDG.128", r"\bLDG\.E\.[A-Z.]*128"),
    ("STG.128", r"\bSTG\.E\.[A-Z.]*128"), ("SYNCS", r"\bSYNCS"), ("F*2", r"\bF(FMA|ADD|MUL)2\b"), ("ST.E", r"\bST\.E\b"),
    ("UCGABAR", r"\bUCGABAR"), ("HMMA", r"\bHMMA"), ("instr", r"^\s*/\*[0-9a-f]{4,}\*/"),
])
kern = None
counts = collections.OrderedDict()
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    for k, p in pats.items():
        if re.search(p, ln):
            counts[kern][k] += 1
dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass juzhen_b200/libjz_b200.so, per kernel (scripts/sass_summary.py); sm_100a only")
print("# " + " | ".join(["kernel"] + list(pats)))
tot = collections.Counter()
for (k, c), name in zip(counts.items(), dem):
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    tot.update(c)
    print(" | ".join([name[:100]] + [str(c[p]) for p in pats]))
print("# TOTAL | " + " | ".join(str(tot[p]) for p in pats))
