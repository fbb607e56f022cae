import json, sys
d=json.load(open(sys.argv[1]))
print("value",round(d["value"],1),"ms",round(d["ms_per_step"],2),"e2e",round(d["e2e"]["value"],1),round(d["e2e"]["ms_per_step"],2), "kp", d["config"]["keypoints_per_volume"], "launches", d["gpu_launches"], "clocks", d["clocks"])
print({k:round(v,3) for k,v in d["stages_ms"].items()})
for k,v in d["kernels"].items(): print(f"{k:14s} ms={v['ms_per_step']:8.3f} n={v['launches']:3d} GB/s={v['alg_gbs']:8.1f} frac={v['frac_hbm']:.3f}")
if d.get("match"): print(d["match"])
if d.get("cpu_baseline"): print(d["cpu_baseline"])
print(d.get("roofline")); print(d.get("dense_pipeline"))
