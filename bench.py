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
  extra : the two paths that SHARD one problem over the N GPUs (strong scaling):
            slab           one 512^3 volume in N z-slabs (halo exchange + all-reduce over NCCL inside the library)
            match_sharded  one 1 M x 1 M enhancedMatch, database sharded N ways, exact re-rank sharded by query

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
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
CPU_BUDGET_S = 150.0   # the reference arm stops starting new steps after this much CPU time (a 512^3 step takes ~1 min)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=512, help="cube edge of the synthetic volume")
    ap.add_argument("--match-n", type=int, default=1000000,
                    help="keypoints per side for the matching legs (BASELINE.json metric: 1 M; 0 = skip)")
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="cube edge of the CPU-baseline volume (0 = the workload's own size: one 512^3 extraction)")
    ap.add_argument("--cpu-match-n", type=int, default=10000, help="keypoints per side of the matcher's CPU baseline (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the strong-scaling legs (slab, match_sharded)")
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


def bind_to_gpu_numa_node(local):
    """Best effort, multi-rank runs only: run this rank on the CPUs of the NUMA node its GPU hangs off BEFORE the pinned
    host buffers are allocated (first touch puts them on that node).  r1/r2 measurements at 8 ranks: the last e2e step's
    H2D took 35 ms on ranks 0-3 and 15-19 ms on ranks 4-7 (profiles/r02_bench_n8_b.json per_rank) — ingest that crosses
    the socket interconnect.  Returns what was done, for the JSON line."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        if all(hasattr(pr, k) for k in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        else:
            q = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                               capture_output=True, text=True, timeout=20).stdout.strip().lower()
            bus = q[-12:] if len(q) >= 12 else q      # nvidia-smi prints an 8-digit domain
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return {"gpu_bus": bus, "node": None, "bound": False, "why": "the platform reports no NUMA node for the GPU"}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if not use:
            return {"gpu_bus": bus, "node": node, "bound": False, "why": "none of the node's CPUs is in this process's cpuset"}
        os.sched_setaffinity(0, use)
        return {"gpu_bus": bus, "node": node, "bound": True, "cpus": len(use)}
    except Exception as ex:   # never fail the bench over placement
        return {"bound": False, "why": f"{type(ex).__name__}: {ex}"}


def reference_checker():
    """The reference's own OpenMP path (oracle/_ref when compiled, else the port) with ALL host threads: torchrun
    exports OMP_NUM_THREADS=1 to its children, which would pin the OpenMP runtime to one core — drop it before the
    library (and with it libgomp) is loaded, and set the reference's own thread count explicitly."""
    os.environ.pop("OMP_NUM_THREADS", None)
    from oracle import ref as O
    chk = O.best()
    if hasattr(chk, "set_threads"):
        chk.set_threads(os.cpu_count() or 1)
    return chk


def time_reference(edge, steps, warmup, seed=0, budget_s=CPU_BUDGET_S):
    """CreateCSIFT3D + KpSiftAlgorithm of the reference on V-blobs(edge): `steps` timed volumes, but no new step is
    started once budget_s of CPU time has been spent (at least one is always timed).
    Returns (Mvoxels/s, mean seconds, checker kind, threads, keypoints, steps timed, stage times of the last step)."""
    synth = importlib.import_module("3dsift_b200.synth")
    chk = reference_checker()
    vol = synth.v_blobs(edge, seed=seed)
    ts, nk, stages = [], 0, {}
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        if i > warmup and time.perf_counter() - t_begin > budget_s:
            break
        t0 = time.perf_counter()
        if chk.kind == "reference":
            sec, nk, stages = chk.time_extract(vol)
        else:
            r = chk.extract(vol, keep_levels=False)
            nk = len(r.keypoints)
            sec = time.perf_counter() - t0
        if i >= warmup:
            ts.append(sec)
    sec = float(np.mean(ts))
    return vol.size / sec / 1e6, sec, chk.kind, chk.threads(), nk, len(ts), {k: round(float(v), 3) for k, v in stages.items()}


def time_reference_match(n, seed=4242):
    """enhancedMatch of the reference's muBruteMatcher on D-synth(n): its own matchTime / totalTime fields
    (Src/cMatcher.cc:197-213)."""
    synth = importlib.import_module("3dsift_b200.synth")
    chk = reference_checker()
    ref, tar, _ = synth.d_synth_pair(n, seed=seed)
    t0 = time.perf_counter()
    r = chk.match(3, ref, tar, 0.85)
    wall = time.perf_counter() - t0
    times = r.get("times")
    total = float(times[2]) if times is not None and times[2] > 0 else wall
    fwd = float(times[0]) if times is not None and times[0] > 0 else None
    return {"value": n * n / total, "unit": "pairs/s", "cores": chk.threads(), "kind": chk.kind, "seconds": total,
            "forward_search_seconds": fwd,
            "sample": f"D-synth {n} x {n}, enhancedMatch thr 0.85, the reference's own totalTime (Src/cMatcher.cc:197-213); "
                      f"1 M x 1 M would take ~{(1e6 / n) ** 2 * total / 3600:.0f} h on these cores (extrapolated by pairs)",
            "matches": int(len(r["pairs"]))}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the SAME configuration as the GPU arm: one whole 512^3 volume per step.  Bounded by time, not by shrinking the
    # volume: at least one step is timed, further ones only while the CPU budget lasts.
    edge = a.cpu_sample or a.size
    steps, warmup = max(1, min(a.steps, 20)), 0
    # a 64^3 volume first: thread pool, page cache and allocator warm (its time is not counted)
    if edge > 64:
        time_reference(64, 1, 0)
    val, sec, kind, threads, nk, done, stages = time_reference(edge, steps, warmup)
    model, ncpu = cpu_info()
    sample = (f"V-blobs {edge}^3" + (" (the workload itself)" if edge == a.size else " (a reduced sample: --cpu-sample)") +
              f", CreateCSIFT3D+KpSiftAlgorithm, {done} timed volume(s) of {sec:.1f} s, "
              f"{threads} OpenMP threads on {ncpu} cores")
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": done, "warmup": warmup,
           "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic",
           "config": {"workload": f"V-blobs {a.size}^3 float32 volume, full extraction", "sample_edge": edge, "sample": sample,
                      "cpu": model, "threads": threads, "stage_seconds": stages},
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
    # stdout carries ONE JSON line: libraries that print there from C (NCCL's version banner) are sent to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 and os.environ.get("S3D_NUMA_BIND", "1") != "0" else {"bound": False, "why": "single rank"}
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"   # NCCL prints its version banner on STDOUT: rank 0 must print one JSON line only
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    s3d = importlib.import_module("3dsift_b200")
    synth = importlib.import_module("3dsift_b200.synth")
    D = importlib.import_module("3dsift_b200.dist")
    api = s3d.api
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

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

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

    # e2e host side.  A step has no host round trip (s3d_run_async only enqueues; the counts stay on the device until
    # s3d_wait), so ONE host thread keeps E2E_DEPTH volumes in flight on private handles/streams: volume i+1 uploading,
    # volume i extracting, volume i-1's results on their way to the host.  S3D_E2E_THREADS > 0 selects the round-1
    # variant instead (blocking KpSiftAlgorithm calls dealt to several host threads).
    _thr_env = os.environ.get("S3D_E2E_THREADS")
    E2E_THREADS = max(0, int(_thr_env)) if _thr_env else 0
    E2E_DEPTH = max(1, int(os.environ.get("S3D_E2E_DEPTH", "2")))

    def e2e_steps_async(k_steps):
        del e2e_step_ms[:]
        t_prev = [time.perf_counter()]
        inflight, kcount = [], 0

        def finish(h, slot):
            h.wait()
            nk = h.num_keypoints()
            h.get_keypoints_async(h_kp[slot].data_ptr(), h_desc[slot].data_ptr())
            h.sync()
            t = h.m_timer
            e2e_split.update(h2d_ms=t["d_h2d"] * 1e3, d2h_ms=t["d_d2h"] * 1e3, device_ms=t["d_TotalTime"] * 1e3)
            now = time.perf_counter()
            e2e_step_ms.append(round((now - t_prev[0]) * 1e3, 2))
            t_prev[0] = now
            h.close()
            return nk

        cur = upload() if k_steps > 0 else None
        for i in range(k_steps):
            nxt = upload() if i + 1 < k_steps else None
            cur.run_async()
            inflight.append((cur, i % n_res))
            if len(inflight) >= E2E_DEPTH:
                kcount = finish(*inflight.pop(0))
            cur = nxt
        while inflight:
            kcount = finish(*inflight.pop(0))
        return kcount

    def e2e_steps(k_steps):
        if E2E_THREADS == 0:
            return e2e_steps_async(k_steps)
        return e2e_steps_threads(k_steps)

    def e2e_steps_threads(k_steps):
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

    def e2e_latency(reps=5):
        """ONE volume at a time through the public host API, nothing overlapped: pinned host volume -> CreateCSIFT3D
        (blocking H2D) -> KpSiftAlgorithm -> GetKeypoints into pinned buffers (blocking D2H).  Median wall ms."""
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            h = s3d.CSIFT3DFactory.CreateCSIFT3D(h_vol, x_dim=n, y_dim=n, z_dim=n, device=local)
            h.KpSiftAlgorithm()
            h.get_keypoints_async(h_kp[0].data_ptr(), h_desc[0].data_ptr())
            h.sync()
            ts.append((time.perf_counter() - t0) * 1e3)
            h.close()
        return float(np.median(ts)), [round(t, 2) for t in ts]

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
    e2e_steps(max(a.warmup, 3 * max(E2E_THREADS, 1)))   # the pipeline reaches its steady state (block cache warm)
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
    lat_ms, lat_all = e2e_latency()
    barrier()

    # ---- strong scaling 1: ONE volume in `world` z-slabs ---------------------------------
    extra = {}
    comm = None
    if not a.no_extra:
        # A failure in a strong-scaling leg must not take the headline line with it: the leg reports the error instead.
        try:
            if world > 1:
                comm = D.nccl_comm()
            else:   # a single-rank communicator of the library's own (no torch.distributed in a one-process run)
                uid = (C.c_ubyte * 128)()
                api.check(L.s3d_comm_unique_id(uid))
                comm = D.NcclComm.__new__(D.NcclComm)
                comm.world, comm.rank, comm._c = 1, 0, C.c_void_p()
                api.check(L.s3d_comm_create(uid, 1, 0, local, C.byref(comm._c)))
            o0, o1 = C.c_int(), C.c_int()
            api.check(L.s3d_slab_bounds(n, world, rank, C.byref(o0), C.byref(o1)))
            own0, own1 = o0.value, o1.value
            d_own, h_own = d_vol[own0:own1], h_vol[own0:own1]
            p_e2e = D._params(dict(device=local))

            def slab_create(resident):
                h = C.c_void_p()
                if resident:
                    api.check(L.s3d_slab_create(comm._c, d_own.data_ptr(), 1, n, n, n, C.byref(p_e2e), C.byref(h)))
                else:
                    api.check(L.s3d_slab_create(comm._c, h_own.data_ptr(), 0, n, n, n, C.byref(p_e2e), C.byref(h)))
                return h

            slab_last = {}

            def slab_steps(k_steps, resident, keep_last=False):
                """k_steps volumes, each extracted by all ranks together, two volumes in flight: s3d_slab_create (allocation +
                copy of the rank's own planes: device-to-device for the resident leg, from pinned host memory for the e2e leg)
                and s3d_slab_execute_async (kernels + NCCL collectives, enqueue only) of volume i+1 are issued while volume i
                runs; s3d_wait, s3d_slab_gather (merge on rank 0 over NCCL) and — e2e leg — the copy of the merged records
                + descriptors into pinned host memory on rank 0 complete volume i."""
                inflight, kk = [], 0

                def finish(h, slot):
                    api.check(L.s3d_wait(h))
                    api.check(L.s3d_slab_gather(comm._c, h, 0, 0))
                    nn = C.c_int()
                    api.check(L.s3d_num_keypoints(h, C.byref(nn)))
                    if rank == 0 and not resident:
                        api.check(L.s3d_get_keypoints_async(h, h_kp[slot].data_ptr(), h_desc[slot].data_ptr()))
                        api.check(L.s3d_sync(h))
                    if keep_last and not inflight:
                        slab_last["h"] = h
                    else:
                        L.s3d_destroy(h)
                    return nn.value

                cur = slab_create(resident) if k_steps > 0 else None
                for i in range(k_steps):
                    nxt = slab_create(resident) if i + 1 < k_steps else None
                    api.check(L.s3d_slab_execute_async(comm._c, cur))
                    inflight.append((cur, i & 1))
                    if len(inflight) >= 2:
                        kk = finish(*inflight.pop(0))
                    cur = nxt
                while inflight:
                    kk = finish(*inflight.pop(0))
                return kk

            slab_steps(max(a.warmup, 3), True)
            sent0 = comm.traffic()[0]
            ksl = max(5, min(a.steps, 20))
            barrier()
            w0 = time.time()
            ev0.record()
            slab_steps(ksl, True, keep_last=True)
            torch.cuda.synchronize()
            ev1.record()
            barrier()
            windows.append((w0, time.time()))
            ms_slab = max_over_ranks(ev0.elapsed_time(ev1) / ksl)
            L.s3d_destroy(slab_last.pop("h"))
            sent_per_step = (comm.traffic()[0] - sent0) / ksl
            # latency of ONE volume, nothing in flight besides it (resident planes): create -> execute -> gather; the
            # per-phase device times are taken from these runs (with two volumes in flight a phase's event span also
            # covers the other volume's kernels)
            lat_slab = []
            for _ in range(5):
                barrier()
                t0 = time.perf_counter()
                slab_steps(1, True, keep_last=True)
                torch.cuda.synchronize()
                lat_slab.append((time.perf_counter() - t0) * 1e3)
                sh = D.SlabShard(slab_last.pop("h"), (n, n, n), rank)
                phases = sh.phases()
                nk_slab = sh.num_keypoints()
                sh.close()
            lat_slab_ms = max_over_ranks(float(np.median(lat_slab)))
            slab_steps(max(a.warmup, 3), False)
            barrier()
            w0 = time.time()
            ev0.record()
            nk_e2e = slab_steps(ksl, False)
            torch.cuda.synchronize()
            ev1.record()
            barrier()
            ms_slab_e2e = max_over_ranks(ev0.elapsed_time(ev1) / ksl)
            windows.append((w0, time.time()))
            ph_all = None
            if world > 1:
                mine = torch.tensor([phases[k_] for k_ in ("normalize", "pyramid", "halo", "sparse", "gather")] + [sent_per_step],
                                    dtype=torch.float64, device="cuda")
                allp = [torch.zeros_like(mine) for _ in range(world)]
                dist.all_gather(allp, mine)
                ph_all = [[round(float(v), 3) for v in x] for x in allp]
            extra["slab"] = {
                "metric": "Mvoxels/s, ONE 512^3 volume extracted by all GPUs together (z-slabs)", "scaling": "strong", "shards": world,
                "value": nvox / (ms_slab * 1e-3) / 1e6, "unit": UNIT, "ms_per_volume": ms_slab, "steps": ksl, "volumes_in_flight": 2,
                "latency_ms_single_volume": lat_slab_ms,
                "e2e": {"value": nvox / (ms_slab_e2e * 1e-3) / 1e6, "unit": UNIT, "ms_per_volume": ms_slab_e2e,
                        "h2d_bytes_per_step": int(vol.nbytes), "d2h_bytes_per_step": int(nk_e2e * (176 + 768 * 4)),
                        "note": "pinned host volume in (every rank uploads its own planes, one volume ahead), merged records + "
                                "descriptors out in pinned host memory on rank 0; CUDA events around the synchronised loop, max over ranks"},
                "keypoints": nk_slab, "equal_to_single_gpu_keypoints": bool(nk_slab == nkp) if rank == 0 else None,
                "phases_ms_rank0": {k_: round(v, 3) for k_, v in phases.items()},
                "phases_per_rank": ph_all, "phases_per_rank_columns": ["normalize ms", "pyramid ms", "halo ms", "sparse ms", "gather ms", "nccl bytes sent per volume"],
                "nccl_bytes_sent_per_volume_rank0": sent_per_step,
                "note": "value = THROUGHPUT with every rank's planes resident in HBM: s3d_slab_create + s3d_slab_execute_async + s3d_wait + "
                        "s3d_slab_gather, two volumes in flight on private streams, CUDA events around the synchronised loop, max over "
                        "ranks; latency_ms_single_volume = one volume at a time; parity with the unsharded run: tests/test_gpu_slab.py "
                        "and scripts/multi_gpu_check.py (bit-equal)"}
        except Exception as ex:   # noqa: BLE001
            extra["slab"] = {"error": f"{type(ex).__name__}: {ex}"}
            print(f"[bench] slab leg failed on rank {rank}: {ex}", file=sys.stderr, flush=True)

    # ---- matching legs (secondary metric): enhancedMatch on HBM-resident descriptor sets ----------------
    match = None
    if a.match_n > 0:
        big = a.match_n >= 200000
        if big:   # generated on the device: numpy would need minutes for 2 x 1 M x 768
            d_ref, d_tar, d_truth = synth.d_synth_pair_device(a.match_n, seed=100)
        else:
            ref, tar, truth = synth.d_synth_pair(a.match_n, seed=100)
            d_ref, d_tar, d_truth = torch.from_numpy(ref).cuda(), torch.from_numpy(tar).cuda(), torch.from_numpy(truth).cuda()
        nr, nt = len(d_ref), len(d_tar)
        I = lambda m: torch.empty(max(m, 1), dtype=torch.int32, device="cuda")
        F = lambda m: torch.empty(max(m, 1), dtype=torch.float32, device="cuda")
        bufs = [I(nr), F(nr), I(nr), F(nr), I(nt), F(nt), I(nt), F(nt), I(nr), I(nr), I(1)]

        def match_step():
            s3d.check(L.s3d_match_device(3, d_ref.data_ptr(), nr, d_tar.data_ptr(), nt, 0.85, *[b.data_ptr() for b in bufs], stream))

        def timed_match(fn):
            for _ in range(1 if big else min(a.warmup, 2)):
                fn()
            msteps = 2 if big else max(1, min(a.steps, 3))
            s3d.match_stats(reset=True)
            barrier()
            w0 = time.time()
            ev0.record()
            for _ in range(msteps):
                fn()
            ev1.record()
            barrier()
            windows.append((w0, time.time()))
            tc_rows, fb_rows = s3d.match_stats()
            return max_over_ranks(ev0.elapsed_time(ev1) / msteps), msteps, tc_rows // msteps, fb_rows // msteps

        ms_match, msteps, tc_rows, fb_rows = timed_match(match_step)
        clk_match = len(windows) - 1
        rev_rows = int((bufs[4] != -1).sum().item())
        npairs = int(bufs[10].item())
        pr_, pt_ = bufs[8][:npairs].long(), bufs[9][:npairs].long()
        true_frac = float((d_truth[pr_] == pt_).float().mean().item()) if npairs else 0.0
        flops = 2.0 * 768 * (nr * nt + rev_rows * nr)
        # sampled exact check: 4096 random forward queries through the exact CUDA-core kernel against the FULL database
        # (float product, sequential double sum — the reference's arithmetic) must give the tensor-core path's results
        g = torch.Generator(device="cpu"); g.manual_seed(7)
        rows = torch.randperm(nr, generator=g)[:min(4096, nr)].cuda()
        qs = d_ref[rows].contiguous()
        d1, d2 = torch.empty(len(rows), dtype=torch.float64, device="cuda"), torch.empty(len(rows), dtype=torch.float64, device="cuda")
        i1, i2 = I(len(rows)), I(len(rows))
        s3d.set_match_path(api.MATCH_EXACT)
        s3d.check(L.s3d_top2_device(qs.data_ptr(), len(rows), d_tar.data_ptr(), nt, 0, None, d1.data_ptr(), i1.data_ptr(),
                                    d2.data_ptr(), i2.data_ptr(), stream))
        s3d.set_match_path(api.MATCH_AUTO)
        torch.cuda.synchronize()
        gI, gD, sI, sD = bufs[0][rows], bufs[1][rows], bufs[2][rows], bufs[3][rows]
        same = (torch.equal(gI.abs(), i1.abs()) and torch.equal(sI, i2) and torch.equal(gD, (2 - 2 * d1).float())
                and torch.equal(sD, (2 - 2 * d2).float()))
        match = {"metric": "match pairs/s (enhancedMatch, thr 0.85)", "n_ref": nr, "n_tar": nt, "ms": ms_match, "steps": msteps,
                 "pairs_per_s": world * nr * nt / (ms_match * 1e-3), "matches": npairs, "true_pair_fraction": true_frac,
                 "sampled_exact_equal": bool(same), "sampled_exact_rows": int(len(rows)),
                 "reverse_rows_searched": rev_rows, "algorithmic_tflops_per_gpu": flops / (ms_match * 1e-3) / 1e12,
                 "scaling": "weak (every GPU matches the same pair of sets on its own; the database-sharded single-problem path is "
                            "extra.match_sharded)",
                 "data": "D-synth(K) generated on the device (torch)" if big else "D-synth(K) (numpy)",
                 "path": "tcgen05 FP16 candidate pass (running top-8 per query row in the epilogue) + exact FP32-product/FP64-sum "
                         "re-rank with guard",
                 "rows_tensor_core": tc_rows, "rows_exact_fallback": fb_rows}
        single_outputs = [b.clone() for b in bufs]

        # ---- strong scaling 2: ONE enhancedMatch, database sharded over the ranks ----------
        if comm is not None:
            try:
                sent0 = comm.traffic()[0]

                def match_sharded_step():
                    s3d.check(L.s3d_match_sharded(comm._c, 3, d_ref.data_ptr(), nr, d_tar.data_ptr(), nt, 0.85,
                                                  *[b.data_ptr() for b in bufs], C.c_void_p(stream or 1)))
                ms_ms, mst, tc2, fb2 = timed_match(match_sharded_step)
                equal = all(torch.equal(x, y) for x, y in zip(bufs[:10], single_outputs[:10])) and int(bufs[10].item()) == npairs
                eq_all = max_over_ranks(0.0 if equal else 1.0) == 0.0
                extra["match_sharded"] = {
                    "metric": "pairs/s, ONE enhancedMatch over all GPUs (database sharded, exact re-rank sharded by query)",
                    "scaling": "strong", "shards": world, "n_ref": nr, "n_tar": nt, "ms": ms_ms, "steps": mst,
                    "pairs_per_s": nr * nt / (ms_ms * 1e-3), "algorithmic_tflops_total": flops / (ms_ms * 1e-3) / 1e12,
                    "equal_to_single_gpu_outputs": bool(eq_all), "rows_tensor_core_rank0": tc2, "rows_exact_fallback_rank0": fb2,
                    "nccl_bytes_sent_per_match_rank0": (comm.traffic()[0] - sent0) / (mst + (1 if big else min(a.warmup, 2))),
                    "clocks": sampler.summary(windows[-1:]),
                    "note": "sets replicated in HBM on every rank; every rank ends with the complete outputs; CUDA events on the "
                            "calling stream, max over ranks"}
            except Exception as ex:   # noqa: BLE001
                extra["match_sharded"] = {"error": f"{type(ex).__name__}: {ex}"}
                print(f"[bench] match_sharded leg failed on rank {rank}: {ex}", file=sys.stderr, flush=True)
        del d_ref, d_tar, bufs, single_outputs
    time.sleep(0.3)
    sampler.stop()

    # ---- reductions over ranks ---------------------------------------------------------------------
    t = torch.tensor([ms_value, ms_e2e, lat_ms], dtype=torch.float64, device="cuda")
    per_rank = None
    if world > 1:
        # per-rank view (value ms, e2e ms, summed kernel ms of the last resident step, H2D ms of the last e2e step) before the max
        mine = torch.tensor([ms_value, ms_e2e, stage.get("d_TotalTime", 0.0) * 1e3, e2e_split.get("h2d_ms", 0.0), lat_ms,
                             float(numa.get("node")) if numa.get("bound") else -1.0], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"ms_value": [round(float(x[0]), 3) for x in allr], "ms_e2e": [round(float(x[1]), 3) for x in allr],
                    "device_ms_last_step": [round(float(x[2]), 3) for x in allr], "h2d_ms_last_e2e_step": [round(float(x[3]), 3) for x in allr],
                    "latency_ms": [round(float(x[4]), 3) for x in allr],
                    "numa_node_bound_to": [int(x[5]) for x in allr], "numa_rank0": numa}
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_value, ms_e2e, lat_ms = float(t[0]), float(t[1]), float(t[2])
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
        try:   # executed warp-instructions per launch of the issue-bound kernels, from the committed ncu capture
            inst_tab = json.load(open(os.path.join(ROOT, "profiles", "instructions.json")))
        except Exception:
            inst_tab = {}
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "B200_PROFILING.md fallback 6650 GB/s"
        kernels = {}
        for name, s in kstats.items():
            per = s["ms"] / max(s["launches"], 1)
            gbs = s["alg_bytes"] / (s["ms"] * 1e-3) / 1e9 if s["ms"] > 0 else 0.0
            kernels[name] = {"ms_per_step": s["ms"], "launches": s["launches"], "avg_launch_ms": per,
                             "alg_bytes": s["alg_bytes"], "alg_gbs": gbs, "frac_hbm": gbs / hbm_peak}
        # the dominant HBM-bound kernel class of the dense pyramid BY TIME, whatever its fraction (r1 judged the choice
        # among the Z / Y / X passes only: the fused X+Y pass takes more time and sits lower)
        dense_names = [k_ for k_ in kernels if k_.startswith("blur") or k_ in ("downsample", "maxabs", "normalize", "detect")]
        roof = None
        cand = [k_ for k_ in dense_names if k_.startswith("blur") and k_ != "blur_generic"]
        if cand:
            top = max(cand, key=lambda k_: kernels[k_]["ms_per_step"])
            kk = kernels[top]
            roof = {"kernel": top, "bound": "hbm", "achieved": kk["alg_gbs"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": kk["frac_hbm"],
                    "traffic": (traffic_tab[top]["dram_bytes_per_step"] / max(kk["launches"], 1)) if top in traffic_tab else None,
                    "alg_bytes_per_launch": kk["alg_bytes"] / max(kk["launches"], 1), "launches": kk["launches"],
                    "traffic_source": traffic_tab.get(top, {}).get("source"), "peak_source": peak_src,
                    "chosen_from": {k_: round(kernels[k_]["ms_per_step"], 3) for k_ in cand},
                    "note": "the dense-pyramid kernel class with the largest summed time in a step; achieved = algorithmic bytes of "
                            "all its launches in a step / their summed CUDA-event time; see profiles/ for ncu dram bytes"}
        sm_clock = (sampler.summary(windows[:1]).get("sm_mhz") or 1965.0) * 1e6
        issue_peak = 148 * 4 * sm_clock   # warp-instructions per second: 4 schedulers per SM, one issue per cycle each
        issue_roof = {}
        for kname in ("describe", "orient", "blur_xy"):
            if kname in kernels and kname in inst_tab and kernels[kname]["ms_per_step"] > 0:
                ach = inst_tab[kname]["warp_instructions_per_step"] / (kernels[kname]["ms_per_step"] * 1e-3)
                issue_roof[kname] = {"bound": "issue", "achieved": ach / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-inst/s",
                                     "frac": ach / issue_peak, "instructions_source": inst_tab[kname].get("source")}
        b_dense = 105.0 * nvox
        dense_ms = sum(kernels[k_]["ms_per_step"] for k_ in dense_names)
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
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e, "host_threads": max(E2E_THREADS, 1), "volumes_in_flight": E2E_DEPTH if E2E_THREADS == 0 else 3 * E2E_THREADS,
                    "h2d_bytes_per_step": int(vol.nbytes),
                    "d2h_bytes_per_step": int(k * (176 + 768 * 4)),
                    "latency_ms_single_volume": lat_ms, "latency_ms_samples_rank0": lat_all,
                    "last_step_split_ms": e2e_split, "host_wall_ms_per_step": list(e2e_step_ms),
                    "note": "value = THROUGHPUT: pinned host volume -> CreateCSIFT3D (H2D on the handle's stream, enqueued one volume "
                            "ahead) -> KpSiftAlgorithm (s3d_run_async: no host round trip inside a step) -> GetKeypoints (D2H of records + "
                            "descriptors into pinned buffers after s3d_wait, under the next volume's kernels); one host thread, "
                            "volumes_in_flight private handles/streams; every volume's H2D and "
                            "D2H complete before the timer stops; host_wall_ms_per_step = completion-to-completion times.  "
                            "latency_ms_single_volume = one volume at a time, nothing overlapped (blocking H2D, extraction, blocking "
                            "D2H), median of 5, max over ranks"},
            "gpu_launches": int(launches),
            "per_rank": per_rank,
            "roofline": roof,
            "issue_roofline": issue_roof or None,
            "dominant_by_time": ({"kernel": "describe", "ms_per_step": kernels["describe"]["ms_per_step"],
                                  "share_of_step": kernels["describe"]["ms_per_step"] / ms_value,
                                  "bound": "issue rate (not HBM: ncu dram throughput 1.5 %; see issue_roofline.describe and "
                                           "profiles/)",
                                  "keypoints_per_s": nkp / (kernels["describe"]["ms_per_step"] * 1e-3)}
                                 if "describe" in kernels and kernels["describe"]["ms_per_step"] > 0 else None),
            "dense_pipeline": {"alg_bytes": b_dense, "ms": dense_ms,
                               "frac_hbm": (b_dense / (dense_ms * 1e-3) / 1e9 / hbm_peak) if dense_ms > 0 else None,
                               "note": "B_dense = 105 bytes/voxel (SURVEY.md §8d) over the summed time of the dense kernels"},
            "stages_ms": {k_: v * 1e3 for k_, v in stage.items()},
            "kernels": kernels,
            "match": match,
            "extra": extra or None,
        }
        if match:
            match["clocks"] = sampler.summary(windows[clk_match:clk_match + 1])   # a 1 M x 1 M search runs into the 1 kW power cap
            match["roofline"] = {"bound": "tensor", "achieved": match["algorithmic_tflops_per_gpu"], "peak": tc_peak, "unit": "TFLOP/s",
                                 "frac": match["algorithmic_tflops_per_gpu"] / tc_peak,
                                 "peak_source": ("MEASURED_PEAKS.json bf16 cuBLAS (sustained)" if peaks else
                                                 "B200_PROFILING.md fallback 1.4 PFLOP/s sustained (1.59 burst)"),
                                 "note": "achieved = 2*768*(rows searched forward + reverse) / whole enhancedMatch time (candidate "
                                         "kernel + re-rank + filters); ncu tensor-pipe utilisation of the candidate kernel is in profiles/"}
        if world == 1 and not a.no_cpu_baseline:
            edge = a.cpu_sample or a.size
            if edge > 64:
                time_reference(64, 1, 0)
            if not a.cpu_sample and edge > 256:
                # the default line must finish within minutes on any host: a 128^3 probe predicts the full-size run (the
                # reference gets ~1.6x faster per voxel from 128^3 to 512^3); above ~2.5 min the baseline falls back to 256^3
                _, probe_s, *_ = time_reference(128, 1, 0)
                if probe_s * (edge / 128.0) ** 3 / 1.6 > 150.0:
                    edge = 256
            val, sec, kind, threads, nk, done, stages = time_reference(edge, 1, 0)
            model, _ = cpu_info()
            out["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "cpu": model, "sample_edge": edge,
                                   "stage_seconds": stages,
                                   "sample": f"V-blobs {edge}^3" + (" (the workload itself)" if edge == a.size else "") +
                                             f", 1 timed volume after a 64^3 warm-up, CreateCSIFT3D+KpSiftAlgorithm ({sec:.2f} s, {nk} keypoints)"}
            if match and a.cpu_match_n > 0:
                match["cpu_baseline"] = time_reference_match(a.cpu_match_n)
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if comm is not None and world == 1:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
