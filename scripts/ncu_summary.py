"""Compact summary of an .ncu-rep (one block per captured launch):  python scripts/ncu_summary.py rep.ncu-rep [--src N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
M = [("dur_us", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("block", "launch__block_size"), ("regs", "launch__registers_per_thread"),
     ("smem_dyn", "launch__shared_mem_per_block_dynamic"), ("occ_achieved%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
     ("sm_thr%", "sm__throughput.avg.pct_of_peak_sustained_elapsed"), ("dram_thr%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
     ("dram_rd", "dram__bytes_read.sum"), ("dram_wr", "dram__bytes_write.sum"), ("l2_thr%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
     ("l1_thr%", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
     ("ipc_active", "smsp__inst_executed.avg.per_cycle_active"), ("issue_active%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
     ("thr/inst", "smsp__thread_inst_executed_per_inst_executed.ratio"),
     ("fma%", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"), ("alu%", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
     ("fp64%", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), ("lsu%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
     ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
     ("smem_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"), ("smem_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
     ("local_ld", "smsp__inst_executed_op_local_ld.sum"), ("local_st", "smsp__inst_executed_op_local_st.sum")]
ST = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0]
    print("==", name)
    out = []
    for lab, key in M:
        if key in idx and r[idx[key]] != "":
            v = r[idx[key]]
            u = units[idx[key]]
            out.append(f"{lab}={v}{'' if u in ('', '%') else ' ' + u}")
    print("  " + "  ".join(out))
    st = sorted(((float(r[idx[h]] or 0), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in ST), reverse=True)
    print("  stalls: " + "  ".join(f"{n}={v:.2f}" for v, n in st[:7]))
if "--src" in sys.argv:
    n = int(sys.argv[sys.argv.index("--src") + 1])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    cur, hdr2, agg = "", None, {}
    for r in csv.reader(io.StringIO(src)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr2 = r
        elif r[0].isdigit() and hdr2 and len(r) >= len(hdr2) - 2:
            try:
                smp = int(r[hdr2.index("# Samples")]); ins = int(r[hdr2.index("Instructions Executed")])
            except (ValueError, IndexError):
                continue
            k = (cur, int(r[0]), r[1].strip())
            a = agg.setdefault(k, [0, 0])
            a[0] += smp; a[1] += ins
    ts = sum(a[0] for a in agg.values()) or 1
    ti = sum(a[1] for a in agg.values()) or 1
    print(f"  top source lines by warp-instructions executed (total {ti}):")
    for (f, ln, txt), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:n]:
        print(f"   inst {a[1] * 100.0 / ti:5.1f}%  samples {a[0] * 100.0 / ts:5.1f}%  {f}:{ln}  {txt[:110]}")
