"""Opcode histogram of one kernel in the product library (cuobjdump -sass): instruction counts per mnemonic, a
cheap check of instruction-count changes before GPU time is spent.  usage: python tools/sass_hist.py <substring> [lib]"""
import re, subprocess, sys, collections
sub = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else "lightplane_b200/csrc/liblightplane_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, hist, total = None, collections.Counter(), 0
for line in txt.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); continue
    if cur and sub in cur:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            hist[m.group(1)] += 1; total += 1
print(sub, "instructions:", total)
print(" ".join(f"{k}:{v}" for k, v in hist.most_common(45)))
