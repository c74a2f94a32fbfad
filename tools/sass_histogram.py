#!/usr/bin/env python
"""Per-kernel SASS opcode census of librnvp_b200.so (cuobjdump -sass): the mnemonics that prove the Blackwell-native paths
(UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk) and those that would
betray a legacy path (HMMA = mma.sync).  Usage: sass_histogram.py [lib.so] > profiles/rNN_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "probaforms_b200/csrc/librnvp_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "HMMA", "FFMA", "MUFU", "LDS", "STS", "LDG", "STG", "RED", "ATOM",
        "SYNCS", "BAR"]
cur, hist, total = None, collections.OrderedDict(), collections.Counter()
for line in out.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", cur).split("(")[0]
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for k in KEYS:
            if op.startswith(k):
                hist[cur][k] += 1
                break
print(f"# SASS opcode census of {lib} (sm_100a), instructions per kernel\n")
print(f"{'kernel':74s} {'instr':>7s} " + " ".join(f"{k:>7s}" for k in KEYS))
for k, h in hist.items():
    print(f"{k[:74]:74s} {total[k]:7d} " + " ".join(f"{h.get(x, 0):7d}" for x in KEYS))
tot = collections.Counter()
for h in hist.values():
    tot.update(h)
print(f"\n{'TOTAL':74s} {sum(total.values()):7d} " + " ".join(f"{tot.get(x, 0):7d}" for x in KEYS))
print("\nHMMA (legacy mma.sync) instructions in the library:", tot.get("HMMA", 0))
