#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native 3DSIFT hot path.

    python bench.py --gpus 1 --steps K --warmup W            # this implementation
    python bench.py --impl reference --steps K --warmup W    # the reference's own OpenMP CPU path

Metric (BASELINE.json): Mvoxels/s of full extraction (CreateCSIFT3D + KpSiftAlgorithm) on a
synthetic 512^3 float32 volume (configs[2] — the configuration the metric is quoted on; it fits one
B200).  One "step" = one whole extraction of one volume.  N > 1 (torchrun): every rank extracts its
own 512^3 volume (independent units, no data-path collective) -> weak scaling; value = all
volumes' voxels / max-over-ranks time.

  value : inputs resident in HBM when the timed region starts (s3d_create_device + s3d_run).
  e2e   : the same through the public host API with HOST buffers: pinned volume -> CreateCSIFT3D
          (H2D inside) -> KpSiftAlgorithm -> GetKeypoints (D2H of records + descriptors).
  match : secondary metric of BASELINE.json — enhancedMatch pairs/s on descriptor sets resident in HBM.

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mvoxels/s extract @512^3"
UNIT = "Mvoxels/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=512, help="cube edge of the synthetic volume")
    ap.add_argument("--match-n", type=int, default=1000000,
                    help="keypoints per side for the matching leg (BASELINE.json metric: 1 M; 0 = skip)")
    ap.add_argument("--cpu-sample", type=int, default=128, help="cube edge of the CPU-baseline sample volume")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="do not bracket kernels with events in the timed steps")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, indices, enabled=True):
        # ONE sampler per job (rank 0), covering every GPU of the job: nvidia-smi queries take driver locks, and a
        # poller per rank measurably slows the other ranks' CUDA calls
        self.indices, self.rows, self.proc, self.enabled = list(indices), [], None, enabled

    def start(self):
        if not self.enabled:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", ",".join(str(i) for i in self.indices)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, windows):
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            if not any(a - 0.05 <= t <= b + 0.05 for a, b in windows):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_info():
    model = ""
    try:
        for l in open("/proc/cpuinfo"):
            if l.startswith("model name"):
                model = l.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return model, os.cpu_count()


def time_reference(edge, steps, warmup, seed=0):
    """The reference's own OpenMP path (oracle/_ref when compiled, else the port) on a bounded
    sample volume; returns (Mvoxels/s, per-step seconds, checker kind, threads, keypoints)."""
    from oracle import ref as O
    synth = importlib.import_module("3dsift_b200.synth")
    chk = O.best()
    vol = synth.v_blobs(edge, seed=seed)
    ts, nk = [], 0
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if chk.kind == "reference":
            sec, nk, _ = chk.time_extract(vol)
        else:
            r = chk.extract(vol, keep_levels=False)
            nk = len(r.keypoints)
            sec = time.perf_counter() - t0
        if i >= warmup:
            ts.append(sec)
    sec = float(np.mean(ts))
    return vol.size / sec / 1e6, sec, chk.kind, chk.threads(), nk


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded: the whole --steps/--warmup run must end within a few minutes on the host cores
    steps, warmup = max(1, min(a.steps, 20)), min(a.warmup, 3)
    val, sec, kind, threads, nk = time_reference(a.cpu_sample, steps, warmup)
    model, ncpu = cpu_info()
    sample = (f"V-blobs {a.cpu_sample}^3 (same generator as the {a.size}^3 workload), CreateCSIFT3D+KpSiftAlgorithm, "
              f"{steps} timed + {warmup} warm-up volumes")
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": steps, "warmup": warmup,
           "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic",
           "config": {"workload": f"V-blobs {a.size}^3 float32 volume, full extraction", "sample": sample, "cpu": model},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "keypoints_per_volume": nk}
    print(json.dumps(out), flush=True)


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"   # NCCL prints its version banner on STDOUT: rank 0 must print one JSON line only
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    s3d = importlib.import_module("3dsift_b200")
    synth = importlib.import_module("3dsift_b200.synth")
    L = s3d.lib()
    s3d.selftest(local)

    n = a.size
    # weak scaling = the SAME per-GPU work on every rank: every rank extracts V-blobs(n, seed 0) from its own pinned host
    # copy (with seed = rank the keypoint count, hence the descriptor time, differed by ~10 % between ranks and the
    # max over ranks measured the heaviest volume instead of the scaling)
    vol = synth.v_blobs(n, seed=0)
    h_vol = torch.from_numpy(vol).pin_memory()
    d_vol = h_vol.cuda(non_blocking=False)
    nvox = vol.size
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(profile):
        sift = s3d.CSIFT3DFactory.CreateCSIFT3D(d_vol, device=local, profile=profile, stream=stream)
        sift.KpSiftAlgorithm()
        return sift

    # pinned result buffers for the e2e leg (sized generously from a first run)
    first = step_resident(False)
    nk0 = first.num_keypoints()
    cap = int(nk0 * 1.5) + 1024
    n_res = 2 * max(3, int(os.environ.get("S3D_E2E_THREADS", "3")))   # two result buffers per e2e host thread
    h_kp = [torch.empty((cap, 176), dtype=torch.uint8).pin_memory() for _ in range(n_res)]
    h_desc = [torch.empty((cap, 768), dtype=torch.float32).pin_memory() for _ in range(n_res)]
    first.close()

    # e2e: the public host API with HOST buffers.  Every step uploads its own 512 MiB volume from pinned
    # memory and reads its keypoints + descriptors back; the upload of step i+1 is enqueued
    # (s3d_create_async, private stream per handle) before step i's extraction so the copy engine
    # works under the kernels — a two-deep software pipeline, all of it inside the timed region.
    def upload():
        return s3d.CSIFT3DFactory.CreateCSIFT3D(h_vol, x_dim=n, y_dim=n, z_dim=n, device=local, async_upload=True)

    e2e_split = {}
    e2e_step_ms = []

    # host threads of the e2e leg: every thread's first upload is exposed (9.7 ms each, one copy engine direction), so
    # concurrency only pays off on longer runs - measured at 10 steps: 19.6 ms/step with 1 thread, 18.9 with 3, 20.8 with 4
    _thr_env = os.environ.get("S3D_E2E_THREADS")
    E2E_THREADS = max(1, int(_thr_env)) if _thr_env else (1 if a.steps < 16 else 2 if a.steps < 32 else 3)

    def e2e_steps(k_steps):
        """k_steps volumes through the public host API, dealt to E2E_THREADS host threads (the C calls release the GIL).
        Each thread keeps three volumes in flight on private handles/streams: volume i+1 uploading (copy engine), volume i
        extracting, volume i-1's records + descriptors on their way into one of the thread's two pinned result buffers
        (s3d_get_keypoints_async, collected with s3d_sync after the next extraction; the last one before the thread ends).
        Several extractions in flight let the GPU fill one volume's kernel tails and its issue-bound descriptor stage
        with another volume's memory-bound pyramid (measured, resident volumes: 18.1 / 17.9 / 17.5 / 16.7 ms per volume with
        1 / 2 / 3 / 4 threads); the handles' buffers come from the library's block cache, so overlapping lifetimes do not
        touch the driver's allocator.  S3D_E2E_THREADS=1, S3D_E2E_ASYNC_D2H=0: the serial loop with a blocking fetch."""
        del e2e_step_ms[:]
        t_start = time.perf_counter()
        async_d2h = os.environ.get("S3D_E2E_ASYNC_D2H", "1") != "0"
        nthreads = min(E2E_THREADS, max(1, k_steps))
        done_at, kcount, errors = [], [0], []
        lock = threading.Lock()

        def worker(tid, n_steps):
            try:
                torch.cuda.set_device(local)
                kp_buf, desc_buf = h_kp[2 * tid:2 * tid + 2], h_desc[2 * tid:2 * tid + 2]

                def finish(h):
                    h.sync()
                    t = h.m_timer
                    with lock:
                        e2e_split.update(h2d_ms=t["d_h2d"] * 1e3, d2h_ms=t["d_d2h"] * 1e3, device_ms=t["d_TotalTime"] * 1e3)
                        done_at.append(time.perf_counter())
                    h.close()

                cur, pend = (upload() if n_steps > 0 else None), None
                for i in range(n_steps):
                    nxt = upload() if i + 1 < n_steps else None
                    cur.KpSiftAlgorithm()
                    if pend is not None:
                        finish(pend)
                    kcount[0] = cur.num_keypoints()
                    cur.get_keypoints_async(kp_buf[i & 1].data_ptr(), desc_buf[i & 1].data_ptr())
                    pend = cur
                    if not async_d2h:
                        finish(pend)
                        pend = None
                    cur = nxt
                if pend is not None:
                    finish(pend)
            except Exception as ex:   # surfaced by the caller: a failed step must fail the bench
                errors.append(ex)

        shares = [k_steps // nthreads + (1 if t < k_steps % nthreads else 0) for t in range(nthreads)]
        if nthreads == 1:
            worker(0, shares[0])
        else:
            th = [threading.Thread(target=worker, args=(t, shares[t])) for t in range(nthreads)]
            for t in th:
                t.start()
            for t in th:
                t.join()
        if errors:
            raise errors[0]
        prev = t_start
        for t in sorted(done_at):   # completion-to-completion times of the volumes
            e2e_step_ms.append(round((t - prev) * 1e3, 2))
            prev = t
        return kcount[0]

    for _ in range(a.warmup):
        step_resident(False).close()

    sampler = ClockSampler(range(world), enabled=(rank == 0))
    sampler.start()
    time.sleep(0.3)
    windows = []

    # ---- value: HBM-resident ----------------------------------------------------------------------
    launches0 = s3d.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.time()
    ev0.record()
    last = None
    for _ in range(a.steps):
        if last is not None:
            last.close()
        last = step_resident(not a.no_profile)
    ev1.record()
    barrier()
    windows.append((w0, time.time()))
    ms_value = ev0.elapsed_time(ev1) / a.steps
    launches = (s3d.launch_count() - launches0) // a.steps
    kstats = last.kernel_stats() if not a.no_profile else {}
    stage = last.m_timer
    nkp = last.num_keypoints()
    n_extre = len(last.extrema()[1])
    last.close()

    # ---- e2e: host buffers, H2D + D2H inside --------------------------------------------------------
    # its own W untimed warm-up steps first: the e2e handles run on private streams, and the first
    # volumes after the resident leg re-home the stream-ordered memory pool's blocks
    e2e_steps(max(a.warmup, 3 * E2E_THREADS))   # every thread reaches its three-volumes-in-flight state (block cache warm)
    barrier()
    w0 = time.time()
    ev0.record()
    t0 = time.perf_counter()
    k = e2e_steps(a.steps)
    torch.cuda.synchronize()
    ev1.record()
    barrier()
    wall_e2e = (time.perf_counter() - t0) / a.steps * 1e3
    windows.append((w0, time.time()))
    ms_e2e = max(ev0.elapsed_time(ev1) / a.steps, wall_e2e)

    # ---- matching leg (secondary metric): enhancedMatch on HBM-resident descriptor sets ----------------
    match = None
    if a.match_n > 0:
        if a.match_n >= 200000:   # generated on the device: numpy would need minutes for 2 x 1 M x 768
            d_ref, d_tar, d_truth = synth.d_synth_pair_device(a.match_n, seed=100 + rank)
        else:
            ref, tar, truth = synth.d_synth_pair(a.match_n, seed=100 + rank)
            d_ref, d_tar, d_truth = torch.from_numpy(ref).cuda(), torch.from_numpy(tar).cuda(), torch.from_numpy(truth).cuda()
        nr, nt = len(d_ref), len(d_tar)
        I = lambda m: torch.empty(max(m, 1), dtype=torch.int32, device="cuda")
        F = lambda m: torch.empty(max(m, 1), dtype=torch.float32, device="cuda")
        bufs = [I(nr), F(nr), I(nr), F(nr), I(nt), F(nt), I(nt), F(nt), I(nr), I(nr), I(1)]

        def match_step():
            s3d.check(L.s3d_match_device(3, d_ref.data_ptr(), nr, d_tar.data_ptr(), nt, 0.85, *[b.data_ptr() for b in bufs], stream))
        big = a.match_n >= 200000
        for _ in range(1 if big else min(a.warmup, 2)):
            match_step()
        msteps = 2 if big else max(1, min(a.steps, 3))
        s3d.match_stats(reset=True)
        barrier()
        w0 = time.time()
        ev0.record()
        for _ in range(msteps):
            match_step()
        ev1.record()
        barrier()
        windows.append((w0, time.time()))
        ms_match = ev0.elapsed_time(ev1) / msteps
        tc_rows, fb_rows = s3d.match_stats()
        rev_rows = int((bufs[4] != -1).sum().item())
        npairs = int(bufs[10].item())
        pr_, pt_ = bufs[8][:npairs].long(), bufs[9][:npairs].long()
        true_frac = float((d_truth[pr_] == pt_).float().mean().item()) if npairs else 0.0
        mt = torch.tensor([ms_match], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(mt, op=dist.ReduceOp.MAX)
        ms_match = float(mt[0])
        flops = 2.0 * 768 * (nr * nt + rev_rows * nr)
        match = {"metric": "match pairs/s (enhancedMatch, thr 0.85)", "n_ref": nr, "n_tar": nt, "ms": ms_match, "steps": msteps,
                 "pairs_per_s": world * nr * nt / (ms_match * 1e-3), "matches": npairs, "true_pair_fraction": true_frac,
                 "reverse_rows_searched": rev_rows, "algorithmic_tflops_per_gpu": flops / (ms_match * 1e-3) / 1e12,
                 "scaling": "weak (every GPU matches its own pair of sets; the database-sharded single-problem path is "
                            "3dsift_b200/dist.py match_sharded, timed by scripts/multi_gpu_check.py)",
                 "data": "D-synth(K) generated on the device (torch)" if big else "D-synth(K) (numpy)",
                 "path": "tcgen05 FP16 candidate pass (running top-8 per query row in the epilogue) + exact FP32-product/FP64-sum "
                         "re-rank with guard",
                 "rows_tensor_core": tc_rows // msteps, "rows_exact_fallback": fb_rows // msteps}
        del d_ref, d_tar, bufs
    time.sleep(0.3)
    sampler.stop()

    # ---- reductions over ranks ---------------------------------------------------------------------
    t = torch.tensor([ms_value, ms_e2e], dtype=torch.float64, device="cuda")
    per_rank = None
    if world > 1:
        # per-rank view (value ms, e2e ms, summed kernel ms of the last resident step) before the max
        mine = torch.tensor([ms_value, ms_e2e, stage.get("d_TotalTime", 0.0) * 1e3], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"ms_value": [round(float(x[0]), 3) for x in allr], "ms_e2e": [round(float(x[1]), 3) for x in allr],
                    "device_ms_last_step": [round(float(x[2]), 3) for x in allr]}
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_value, ms_e2e = float(t[0]), float(t[1])
    value = world * nvox / (ms_value * 1e-3) / 1e6
    e2e = world * nvox / (ms_e2e * 1e-3) / 1e6

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        tc_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
        try:   # dram bytes per launch from the committed ncu capture of the same command (profiles/)
            traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            traffic_tab = {}
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "B200_PROFILING.md fallback 6650 GB/s"
        kernels = {}
        for name, s in kstats.items():
            per = s["ms"] / max(s["launches"], 1)
            gbs = s["alg_bytes"] / (s["ms"] * 1e-3) / 1e9 if s["ms"] > 0 else 0.0
            kernels[name] = {"ms_per_step": s["ms"], "launches": s["launches"], "avg_launch_ms": per,
                             "alg_bytes": s["alg_bytes"], "alg_gbs": gbs, "frac_hbm": gbs / hbm_peak}
        # the dominant HBM-bound kernel: the DoG-fused Z pass (largest dense launches)
        roof = None
        dense = [k for k in ("blur_z_dog", "blur_y", "blur_x") if k in kernels]
        if dense:
            top = max(dense, key=lambda k: kernels[k]["ms_per_step"])
            kk = kernels[top]
            roof = {"kernel": top, "bound": "hbm", "achieved": kk["alg_gbs"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": kk["frac_hbm"],
                    "traffic": (traffic_tab[top]["dram_bytes_per_step"] / max(kk["launches"], 1)) if top in traffic_tab else None,
                    "alg_bytes_per_launch": kk["alg_bytes"] / max(kk["launches"], 1), "launches": kk["launches"],
                    "traffic_source": traffic_tab.get(top, {}).get("source"), "peak_source": peak_src,
                    "note": "achieved = algorithmic bytes of all launches of this kernel class in a step / their summed "
                            "CUDA-event time; see profiles/ for ncu dram bytes"}
        b_dense = 105.0 * nvox
        dense_ms = sum(kernels[k]["ms_per_step"] for k in kernels if k.startswith("blur") or k in ("downsample", "maxabs", "normalize", "detect"))
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"V-blobs {n}^3 float32 volume per GPU (BASELINE.json configs[2] size), full extraction: "
                                   "normalise + Gaussian pyramid + DoG + detection + orientation + 768-d descriptors",
                       "volumes_per_step": world, "per_rank_input": "every rank: V-blobs(n, seed 0), own pinned host copy", "keypoints_per_volume": nkp, "detections_per_volume": n_extre,
                       "l2": f"inputs ({vol.nbytes >> 20} MiB/volume) are larger than L2; no flush needed",
                       "parallelism": f"dp{world} (one volume per GPU, no collective on the data path)"},
            "clocks": sampler.summary(windows[:2]),   # the two extraction legs (value, e2e)
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e, "host_threads": E2E_THREADS, "h2d_bytes_per_step": int(vol.nbytes),
                    "d2h_bytes_per_step": int(k * (176 + 768 * 4)),
                    "last_step_split_ms": e2e_split, "host_wall_ms_per_step": list(e2e_step_ms),
                    "note": "pinned host volume -> CreateCSIFT3D (H2D on the handle's stream, enqueued one volume ahead) -> "
                            "KpSiftAlgorithm -> GetKeypoints (D2H of records + descriptors into pinned buffers, enqueued with "
                            "s3d_get_keypoints_async and collected with s3d_sync after the next extraction); the steps are dealt to "
                            "host_threads threads, each with its own handles, streams and result buffers; every volume's H2D and "
                            "D2H complete before the timer stops; host_wall_ms_per_step = completion-to-completion times"},
            "gpu_launches": int(launches),
            "per_rank": per_rank,
            "roofline": roof,
            "dominant_by_time": ({"kernel": "describe", "ms_per_step": kernels["describe"]["ms_per_step"],
                                  "share_of_step": kernels["describe"]["ms_per_step"] / ms_value,
                                  "bound": "issue rate (not HBM: ncu dram throughput 1.5 %, issue slots 82 % busy; "
                                           "profiles/r01_ncu_describe_*.txt)",
                                  "keypoints_per_s": nkp / (kernels["describe"]["ms_per_step"] * 1e-3)}
                                 if "describe" in kernels and kernels["describe"]["ms_per_step"] > 0 else None),
            "dense_pipeline": {"alg_bytes": b_dense, "ms": dense_ms,
                               "frac_hbm": (b_dense / (dense_ms * 1e-3) / 1e9 / hbm_peak) if dense_ms > 0 else None,
                               "note": "B_dense = 105 bytes/voxel (SURVEY.md §8d) over the summed time of the dense kernels"},
            "stages_ms": {k: v * 1e3 for k, v in stage.items()},
            "kernels": kernels,
            "match": match,
        }
        if match:
            match["clocks"] = sampler.summary(windows[2:])   # a 1 M x 1 M search runs into the 1 kW power cap
            match["roofline"] = {"bound": "tensor", "achieved": match["algorithmic_tflops_per_gpu"], "peak": tc_peak, "unit": "TFLOP/s",
                                 "frac": match["algorithmic_tflops_per_gpu"] / tc_peak,
                                 "peak_source": ("MEASURED_PEAKS.json bf16 cuBLAS (sustained)" if peaks else
                                                 "B200_PROFILING.md fallback 1.4 PFLOP/s sustained (1.59 burst)"),
                                 "note": "achieved = 2*768*(rows searched forward + reverse) / whole enhancedMatch time (candidate "
                                         "kernel + re-rank + filters); ncu tensor-pipe utilisation of the candidate kernel is in profiles/"}
        if world == 1 and not a.no_cpu_baseline:
            val, sec, kind, threads, nk = time_reference(a.cpu_sample, 6, 1)
            model, _ = cpu_info()
            out["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "cpu": model,
                                   "sample": f"V-blobs {a.cpu_sample}^3, 6 timed + 1 warm-up volumes, CreateCSIFT3D+KpSiftAlgorithm "
                                             f"({sec:.2f} s each, {nk} keypoints); the 512^3 run would take minutes"}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
