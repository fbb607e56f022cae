"""Instruction-level view of one kernel in an .ncu-rep (captured with --import-source on):
    python scripts/sass_profile.py rep.ncu-rep [--top N] [--regions]
Prints executed warp-instructions by opcode, shared-memory wavefronts by opcode, and (--regions) the SASS
listing collapsed into runs of equal execution count (= loop nests), with each run's share of the total."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None
ins = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) >= len(hdr) - 2 and r[0].startswith("0x"):
        ins.append(r)
H = {h: i for i, h in enumerate(hdr)}
def col(r, name):
    try:
        return float(r[H[name]])
    except Exception:
        return 0.0
tot = sum(col(r, "Instructions Executed") for r in ins)
tsm = sum(col(r, "# Samples") for r in ins)
print(f"SASS instructions: {len(ins)}, executed warp-instructions: {tot:.4g}, samples: {tsm:.0f}")
byop = collections.Counter(); smp = collections.Counter(); wav = collections.Counter(); ideal = collections.Counter(); thr = collections.Counter()
for r in ins:
    s = r[H["Source"]].strip()
    parts = s.split()
    op = parts[1] if parts and parts[0].startswith("@") else (parts[0] if parts else "?")
    op = op.rstrip(";")
    base = ".".join(op.split(".")[:2]) if op.startswith(("ATOMS", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "LDGSTS", "RED", "ATOM")) else op.split(".")[0]
    e = col(r, "Instructions Executed")
    byop[base] += e; smp[base] += col(r, "# Samples")
    thr[base] += col(r, "Thread Instructions Executed")
    wav[base] += col(r, "L1 Wavefronts Shared"); ideal[base] += col(r, "L1 Wavefronts Shared Ideal")
print(f"{'opcode':<14}{'inst %':>8}{'samples %':>10}{'thr/inst':>9}{'smem wavefronts':>17}{'ideal':>14}")
for op, e in byop.most_common(top):
    print(f"{op:<14}{100*e/tot:8.2f}{100*smp[op]/max(tsm,1):10.2f}{thr[op]/max(e,1):9.1f}{wav[op]:17.4g}{ideal[op]:14.4g}")
print(f"shared wavefronts total {sum(wav.values()):.4g} ideal {sum(ideal.values()):.4g}")
if "--regions" in sys.argv:
    runs = []
    for i, r in enumerate(ins):
        e = col(r, "Instructions Executed")
        if runs and abs(runs[-1][2] - e) <= 0.02 * max(e, runs[-1][2]):
            runs[-1][1] = i; runs[-1][3] += e; runs[-1][4] += col(r, "# Samples")
        else:
            runs.append([i, i, e, e, col(r, "# Samples")])
    print("regions (first..last SASS index, count, per-inst executions, share of instructions, share of samples):")
    for a, b, e, s, sm in runs:
        if s / tot >= 0.004:
            print(f"  [{a:5d}..{b:5d}] n={b-a+1:4d} exec/inst={e:11.4g} inst%={100*s/tot:6.2f} samp%={100*sm/max(tsm,1):6.2f}   {ins[a][H['Source']].strip()[:60]}")
