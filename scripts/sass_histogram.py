"""Per-kernel SASS opcode histogram of the shipped library (no GPU needed):
    python scripts/sass_histogram.py [3dsift_b200/lib/libsift3d_b200.so] > profiles/rNN_sass_histogram.txt
For every kernel: instruction count and the counts of the opcodes that prove (or disprove) a Blackwell-native path —
UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG/UBLKCP (TMA), LDGSTS (cp.async), HMMA (legacy mma.sync),
ATOMS/RED/ATOMG, FFMA/FMUL/FADD/DADD, MUFU, BAR, SHFL."""
import collections, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "3dsift_b200/lib/libsift3d_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "HMMA", "ATOMS", "ATOMG", "RED",
        "LDG", "STG", "LDS", "STS", "FFMA", "FMUL", "FADD", "DADD", "DFMA", "MUFU", "BAR", "SHFL", "MATCH", "VOTE", "BRA"]
cur, tab = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        tab[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        op = m.group(1)
        tab[cur]["_total"] += 1
        for k in KEYS:
            if op == k or (k in ("UTCHMMA", "UTCQMMA", "UTCIMMA") and op.startswith(k)):
                tab[cur][k] += 1
print(f"# SASS opcode histogram of {so} (cuobjdump -sass; sm_100a)")
print(f"{'kernel':<58}{'instr':>7} " + " ".join(f"{k:>7}" for k in KEYS))
for name, c in tab.items():
    short = name.replace("s3d::", "").replace("void ", "")[:57]
    print(f"{short:<58}{c['_total']:>7} " + " ".join(f"{c[k]:>7}" if c[k] else f"{'.':>7}" for k in KEYS))
