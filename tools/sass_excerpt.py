"""profiles/rNN_sass_excerpt.txt: per-kernel SASS evidence (tcgen05 / TMEM / TMA mnemonics) from the shipped library.
   python tools/sass_excerpt.py > profiles/r02_sass_excerpt.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "unbiased-teacher-v2_b200", "lib", "libut2_sm100.so")
WANT = ("conv_fwd_kernel", "conv3x3_halo_kernel", "conv_wgrad_kernel", "stem_tc_kernel", "stem_pool_tc_kernel")
KEYS = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "REDG", "RED.E",
        "STG.E.ENL2.256", "LDS", "STS", "HMMA", "LDG", "ATOMG")
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, per, lines = None, collections.OrderedDict(), {}
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = name if any(w in name for w in WANT) else None
        if cur:
            per[cur] = collections.Counter()
            lines[cur] = []
        continue
    if cur and "/*" in ln and ";" in ln:
        ins = ln.split("*/")[1].split(";")[0].strip() if "*/" in ln else ""
        if not ins:
            continue
        lines[cur].append(ins)
        for k in KEYS:
            if re.search(r"(^|\s|@!?U?P\d\s)" + re.escape(k), ins):
                per[cur][k] += 1
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a) — tensor-core / TMEM / TMA mnemonics per kernel")
for name, c in per.items():
    short = name.replace('(anonymous namespace)::', '').split('(')[0]
    print(f"\n== {short}   ({len(lines[name])} SASS instructions)")
    print("   " + "  ".join(f"{k}x{v}" for k, v in c.items() if v))
    shown = 0
    for ins in lines[name]:
        if any(k in ins for k in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "UTMASTG")) and shown < 10:
            print("      " + ins)
            shown += 1
