#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA (tcgen05.mma),
LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA), plus HMMA (legacy mma.sync - must be 0) and MUFU.EX2.
    python tools/sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt"""
import collections, re, subprocess, sys, os
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "touchnet_b200", "lib", "libtouchnet_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = {"UTCHMMA": r"\bUTC[A-Z]*MMA", "UTCHMMA.2CTA": r"\bUTC[A-Z]*MMA\.2CTA", "LDTM": r"\bLDTM", "STTM": r"\bSTTM", "UTMALDG": r"\bUTMALDG",
        "UTMASTG": r"\bUTMASTG", "UBLKCP": r"\bUBLKCP", "HMMA(legacy)": r"\bHMMA", "MUFU.EX2": r"\bMUFU\.EX2", "RED/ATOM": r"\b(RED|ATOMG|ATOM)\b"}
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void ", "")
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for k, p in pats.items():
        if re.search(p, line):
            counts[cur][k] += 1
keys = list(pats)
print(f"# {os.path.basename(lib)}: SASS mnemonic counts per kernel (cuobjdump -sass); tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG")
print("kernel," + ",".join(keys))
tot = collections.Counter()
for k, c in counts.items():
    if sum(c.values()) == 0:
        continue
    print(k + "," + ",".join(str(c[x]) for x in keys))
    tot.update(c)
print("TOTAL," + ",".join(str(tot[x]) for x in keys))
