"""Static SASS opcode histogram per kernel of the shipped library (not a test): python tests/tools/sass_opcodes.py > profiles/r2_sass_opcodes.md"""
import collections, re, subprocess, sys
LIB = "circuitmap_b200/libcircuitmap_b200.so"
COLS = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "DMMA", "HMMA", "REDUX", "LDGSTS",
        "SHFL", "BAR", "ATOMS", "RED", "MUFU", "DFMA", "STL", "LDL"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
dem = {}
kern, counts, total = None, {}, {}
pat = re.compile(r"^\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)")
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    m = pat.match(line)
    if m and kern:
        counts[kern][m.group(1)] += 1
names = list(counts)
d = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
print("# SASS opcode histogram of `%s` (end of round 2, `cuobjdump -sass`, `tests/tools/sass_opcodes.py`)\n" % LIB)
print("Static instruction counts per kernel (every instruction of the kernel, sub-opcodes folded: `DMMA.8x8x4` counts as `DMMA`).  `UTCHMMA` =\n"
      "`tcgen05.mma kind::f16/tf32`, `LDTM` = `tcgen05.ld`, `UBLKCP` = `cp.async.bulk` (1-D bulk copy, mbarrier completion; there is no tensor-map\n"
      "`UTMALDG` / `UTMASTG` in the library), `SYNCS` = mbarrier ops, `DMMA` = `mma.sync.m8n8k4.f64`, `REDUX` = `redux.sync` (exact-0/1 counts of the\n"
      "chain sweep), `LDGSTS` = `cp.async`, `STL` / `LDL` = local-memory (spill) stores / loads.\n")
print("| kernel | total | " + " | ".join(COLS) + " |")
print("|---|---|" + "---|" * len(COLS))
for n, dn in sorted(zip(names, d), key=lambda t: -sum(counts[t[0]].values())):
    c = counts[n]
    print("| `%s` | %d | %s |" % (dn[:90], sum(c.values()), " | ".join(str(c[k]) for k in COLS)))
