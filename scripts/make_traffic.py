"""profiles/traffic.json from an ncu capture of the blur kernels of one 512^3 step (metrics
dram__bytes_read.sum,dram__bytes_write.sum are enough; see scripts/gpu_traffic.sh):
    python scripts/make_traffic.py gpurun_out/traffic_blur.ncu-rep
Per bench kernel class: mean dram__bytes_read.sum + dram__bytes_write.sum per launch (what bench.py reports as roofline.traffic)."""
import csv, io, json, os, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
agg = {}
prev = ""
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    # a level is blurred as (fused XY | X, Y) then Z: a march that follows blur_x_kernel is the Y pass, any other the Z pass
    # (bench.py's class blur_z_dog holds every Z pass, with or without the fused DoG)
    if "blur_march" in name:
        cls = "blur_y" if prev == "blur_x" else "blur_z_dog"
    elif "blur_xy" in name:
        cls = "blur_xy"
    elif "blur_x_kernel" in name:
        cls = "blur_x"
    else:
        continue
    prev = cls
    b = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) + to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
    a = agg.setdefault(cls, [0, 0.0])
    a[0] += 1; a[1] += b
out = {k: {"launches": n, "dram_bytes_per_launch": tot / n, "dram_bytes_per_step": tot,
           "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, all launches of one 512^3 extraction ({os.path.basename(rep)})"} for k, (n, tot) in agg.items()}
dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
