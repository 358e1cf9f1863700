#!/usr/bin/env python
"""bench.py - particle-updates/s of the tree-gravity + SIDM hot path (BASELINE.json metric).

Workload (BASELINE.json configs[2], the one the metric is quoted on; it fits one B200):
isolated NFW halo of parameter.txt:6-11, N = 1e7 equal-mass particles, sigma/m = 1 cm^2/g,
eps = 0.3 kpc, ErrTolTheta 0.5, relative opening criterion alpha = 0.005, DesNumNgb 30 +- 2,
every particle active on a fixed step (all-active steps).  Synthetic seeded
ICs (sidm_b200/ic.py), the reference's own parameter values.

A "step" = compute_accelerations(0) (predict + tree build + walk + sidm + ensure_neighbours,
accel.c:27-132) for all N particles followed by advance() (predict.c:245) so that the next
step sees moved particles.  value = N * K / time with the particle state resident in HBM;
e2e = the same through the drop-in boundary: host array-of-structs (124-byte particle_data)
-> b200_upload -> b200_compute_accelerations -> b200_download, copies inside the timed region.

`--impl reference` times the UNMODIFIED reference (oracle/_ref/libsidmref_fast.so, built from
/root/reference by oracle/Makefile) on the host cores: P forked single-rank copies each hold
the full N-particle system and advance a disjoint random sample of the active list through
the reference's own compute_accelerations(0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "sidm-nbody_b200"))

METRIC = "particle-updates/s (tree gravity+SIDM DSMC), N=1e7 halo, 1/2/4/8 B200"
UNIT = "particle-updates/s"
SIGMA_CM2_G = 1.0
DT = 1.0e-4                # internal time units (0.978 Gyr): the G*rho / SIDM step limit of timestep.c:247-265 at the
                           # centre of this halo at N=1e7, so that an all-active step is one the reference would take
WORKLOAD = "NFW halo N={n:.0e} (rho0=1.49e-4, rs=11.14 kpc, rmax=100 rs, seed 3), sigma/m=1 cm^2/g, eps=0.3 kpc, all-active steps dt=1e-4"


def make_ic(n):
    from sidm_b200 import ic
    return ic.nfw(n, seed=3)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region"""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ reference arm

def _ref_worker(conn, n, pos, vel, mass, ids, sample, sigma_int, dt):
    """one single-rank copy of the unmodified reference; advances `sample` through compute_accelerations(0)"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refdrv
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)                                  # the reference prints progress lines to stdout
    try:
        R = refdrv.Reference("fast")
        R.setup(n, CrossSectionInternal=sigma_int, TreeUpdateFrequency=0.1, BufferSizeMB=100)
        R.init_rand(55)
        R.set_particles(pos, vel, mass, ids)
        R.treebuild()
        h = np.zeros(n, np.float32)
        for i in sample:
            h[i] = np.sqrt(R.ngb_treefind(pos[i], 30))
        R.set("HSML", h)
        R.set("NGB", np.full(n, 30, np.int32))           # only the sample is ever out of range
        R.all_active(0.0, dt / 2)                        # CurrentTime 0, prediction time dt/2
        R.getvmax()
        conn.send(("ready", 0.0))
        while True:
            cmd = conn.recv()
            if cmd == "stop":
                break
            R.set_active(sample)
            t0 = time.perf_counter()
            R.compute_accelerations(0)
            t1 = time.perf_counter()
            conn.send(("done", t1 - t0))
    except Exception as e:  # pragma: no cover
        conn.send(("error", repr(e)))


def run_reference_sample(n, nproc, per_proc, warmup, steps, ic_data=None):
    """returns (updates_per_s, cores, sample_description, seconds_per_step)"""
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refdrv
    if not refdrv.available("fast"):
        return None
    from sidm_b200 import ic
    pos, vel, mass, ids = ic_data if ic_data is not None else make_ic(n)
    sigma_int = ic.cross_section_internal(SIGMA_CM2_G)
    rng = np.random.default_rng(12345)
    allsample = rng.choice(n, size=min(n, nproc * per_proc), replace=False).astype(np.int32)
    parts = np.array_split(allsample, nproc)
    ctx = mp.get_context("fork")
    procs = []
    for p in range(nproc):
        a, b = ctx.Pipe()
        pr = ctx.Process(target=_ref_worker, args=(b, n, pos, vel, mass, ids, np.sort(parts[p]), sigma_int, DT), daemon=True)
        pr.start()
        procs.append((pr, a))
    for pr, a in procs:
        tag, v = a.recv()
        if tag != "ready":
            raise RuntimeError(f"reference worker failed: {v}")
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        for pr, a in procs:
            a.send("step")
        for pr, a in procs:
            tag, v = a.recv()
            if tag != "done":
                raise RuntimeError(f"reference worker failed: {v}")
        t1 = time.perf_counter()
        if s >= warmup:
            times.append(t1 - t0)
    for pr, a in procs:
        a.send("stop")
    for pr, a in procs:
        pr.join(timeout=10)
    tstep = float(np.mean(times))
    desc = (f"{len(allsample)} of {n} particles active per step ({nproc} forked single-rank copies of the unmodified reference x "
            f"{len(parts[0])} targets, each holding the full {n}-particle tree; shipped TreeUpdateFrequency 0.1 so the tree build is amortised as in the reference)")
    return len(allsample) / tstep, nproc, desc, tstep


def host_parallelism(n):
    cores = os.cpu_count() or 1
    try:
        import psutil
        avail = psutil.virtual_memory().available
        per = n * (124 + 0.8 * 136 + 16) + 300e6        # P[], nodes, links, comm buffer + ngb lists
        cores = max(1, min(cores, int(avail * 0.7 / per)))
    except Exception:
        pass
    return max(1, min(cores, 64))


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.n
    nproc = args.ref_procs or host_parallelism(n)
    t0 = time.time()
    r = run_reference_sample(n, nproc, args.ref_sample, args.warmup, args.steps)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libsidmref_fast.so not built (needs /root/reference at build time)"}))
        return
    value, cores, desc, tstep = r
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": tstep * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD.format(n=n), "particles": n, "parallelism": f"host cpu x{cores}"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.time() - t0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ B200 arm

ROOF_STEPS = 2


def main_b200(args):
    import torch
    import torch.distributed as dist
    from sidm_b200 import HotPath, capi, ic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 arm has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.n
    pos, vel, mass, ids = make_ic(n)
    sigma_int = ic.cross_section_internal(SIGMA_CM2_G)
    hp = HotPath(n, device=local, CrossSectionInternal=sigma_int, Seed=55)

    from sidm_b200.multi import Sharder
    sh = Sharder(hp, world, rank)

    # ---- set-up (untimed): start-up forces and smoothing lengths, as init.c:120-180
    hp.set_particles(pos, vel, mass, ids)
    hp.predict_collisionless_only(0.0)
    hp.force_treebuild()
    hp.setup_smoothinglengths_sidm(30)
    vmax = hp.getvmax()
    sh.compute_accelerations(1, time=0.0, vmax=vmax)      # BH criterion (OldAcc = 0) -> OldAcc

    tcur = 0.0

    def step():
        nonlocal tcur
        t = tcur + DT / 2
        sh.compute_accelerations(0, time=t, vmax=vmax)
        hp.advance(time=t)
        tcur += DT

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    c0 = hp.counters()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    walk_ms, build_ms, sidm_ms, ens_ms = [], [], [], []
    list_nodes = list_parts = ntarg = num_lists = 0
    inter_p = inter_n = 0
    scat = rep_it = rep_n = cand = 0
    ev0.record()
    for _ in range(args.steps):
        step()
        c = hp.counters()
        walk_ms.append(c.ms_walk); build_ms.append(c.ms_build); sidm_ms.append(c.ms_sidm); ens_ms.append(c.ms_ensure)
        list_nodes += c.list_nodes; list_parts += c.list_parts; ntarg += c.num_targets; num_lists += c.num_lists
        inter_p += c.part_interactions; inter_n += c.node_interactions
        scat += c.sct_scattered; rep_it += c.ensure_iterations; rep_n += c.ensure_repaired; cand += c.ngb_candidates
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        tt = torch.tensor([ms], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    clk = clocks.stop() if rank == 0 else None
    c1 = hp.counters()
    launches = (c1.kernel_launches - c0.kernel_launches)
    value = n * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (the tree walk), SURVEY.md 8d bytes formula.  In the timed
    # steps the walk shares the GPU with the SIDM chain (two streams), so its event-bracketed duration
    # there is not the kernel's own; the kernel duration is taken live from ROOF_STEPS further steps of
    # the same run with the two phases in sequence (b200_set_option("overlap", 0)).
    peak, peak_src = peaks()
    wms_overlapped = float(np.mean(walk_ms))
    hp.set_option("overlap", 0)
    iso = []
    for _ in range(ROOF_STEPS):
        step()
        iso.append(hp.counters().ms_walk)
    hp.set_option("overlap", 1)
    wms = float(np.mean(iso))
    a_per_launch = ntarg / args.steps
    lists = num_lists / args.steps                       # interaction lists per launch (one per warp of 32 targets)
    i_n = list_nodes / max(1, args.steps) / lists
    i_p = list_parts / max(1, args.steps) / lists
    walk_bytes = a_per_launch * 32 + lists * (48 * i_n + 16 * i_p)
    achieved = walk_bytes / (wms * 1e-3) / 1e9
    flops = (inter_n * 70.0 + inter_p * 20.0) / args.steps
    traffic = None
    tj = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tj) and world == 1:
        t = json.load(open(tj))
        if int(t.get("particles", 0)) == n:
            traffic = t["traffic_bytes_per_launch"]      # ncu dram read+write bytes of one k_walk launch (profiles/)
    roofline = {"bound": "hbm", "kernel": "k_walk", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "ms_per_launch": wms,
                "algorithmic_bytes_per_launch": walk_bytes, "I_n_per_list": i_n, "I_p_per_list": i_p, "targets_per_list": a_per_launch / lists,
                "interactions_per_target": {"node": inter_n / max(1, ntarg), "particle": inter_p / max(1, ntarg)},
                "fp32_tflops_est": flops / (wms * 1e-3) / 1e12, "ms_per_launch_while_sidm_overlaps": wms_overlapped,
                "note": "walk is FP32-issue bound, not HBM bound: see DESIGN.md section 5"}
    phases = {"build_ms": float(np.mean(build_ms)), "walk_ms": wms, "sidm_ms": float(np.mean(sidm_ms)), "ensure_ms": float(np.mean(ens_ms)),
              "scatterings_per_step": scat / args.steps, "ensure_passes_per_step": rep_it / args.steps,
              "ensure_repaired_per_step": rep_n / args.steps, "ngb_candidates_per_search": cand / max(1, args.steps) / (n + rep_n / args.steps)}

    # ---- end to end through the drop-in boundary (host AoS in pinned memory)
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(hp, sh, n, mass, ids, vmax, tcur, args, world, rank, step)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            nproc = args.ref_procs or host_parallelism(n)
            r = run_reference_sample(n, nproc, max(500, args.ref_sample // 4), 1, 1, (pos, vel, mass, ids))
            if r is not None:
                cpu = {"value": r[0], "unit": UNIT, "cores": r[1], "kind": "reference", "sample": r[2]}
        except Exception as e:  # pragma: no cover
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"failed: {e!r}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": WORKLOAD.format(n=n), "particles": n, "parallelism": sh.describe(),
                           "l2": "inputs larger than L2 (>=1.2 GB of particle+node records per step vs 126 MB L2)",
                           "step": "compute_accelerations(0) + advance(), all particles active"},
                "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "phases": phases}
        print(json.dumps(line))
    hp.close()
    if world > 1:
        dist.destroy_process_group()


def run_e2e(hp, sh, n, mass, ids, vmax, tcur, args, world, rank, step_fn):
    """host AoS -> upload -> compute_accelerations(0) -> download, every step, copies timed.
    The inputs are successive states of the run (one fresh host array per step), captured untimed."""
    import torch
    from sidm_b200 import capi
    nsnap = max(2, min(args.steps, 3))
    snaps, times = [], []
    t = tcur
    for s in range(nsnap):
        a = np.zeros(n, capi.PARTICLE_DTYPE)
        velh = hp.peek("velh", np.float32, (n, 4))
        a["Pos"] = hp.peek("pos0", np.float32, (n, 3)); a["PosPred"] = a["Pos"]
        a["Vel"] = velh[:, :3]; a["VelPred"] = a["Vel"]; a["HsmlVelDisp"] = velh[:, 3]
        a["Mass"] = mass; a["ID"] = ids; a["Type"] = 1
        a["CurrentTime"] = hp.peek("curtime", np.float32, (n,))
        a["Accel"], a["OldAcc"], a["NgbVelDisp"] = hp.get("Accel", "OldAcc", "NgbVelDisp")
        snaps.append(a); times.append(t + DT / 2)
        step_fn()
        t += DT
    out = np.zeros(n, capi.PARTICLE_DTYPE)
    rt = torch.cuda.cudart()
    for a in snaps + [out]:
        rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)
    hp.bind_particles(snaps[0], pin=False)

    def one(k):
        hp.bind_particles(snaps[k % nsnap], pin=False)
        sh.upload()                                  # own rows over PCIe (+ NVLink all-gather when sharded)
        sh.compute_accelerations(0, time=times[k % nsnap], vmax=vmax)
        sh.download(into=out)
        return int(hp.counters().sct_scattered)

    steps = max(2, min(args.steps, 5))
    one(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(steps):
        one(k)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor([dt], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    for a in snaps + [out]:
        rt.cudaHostUnregister(a.ctypes.data)
    rows = sh.rows()[1] if world > 1 else n
    return {"value": n * steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(rows * snaps[0].itemsize * world),
            "d2h_bytes_per_step": int(rows * out.itemsize * world),
            "steps": steps, "ms_per_step": dt / steps * 1e3,
            "api": ("b200_bind_particles + b200_upload + b200_compute_accelerations(0) + b200_download_to on pinned 124-byte particle_data arrays (successive states of the run)"
                    if world == 1 else "b200_upload_shard (own rows over PCIe + NVLink all-gather) + b200_compute_accelerations(0) + b200_download_shard, pinned 124-byte particle_data rows per rank")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=lambda s: int(float(s)), default=10_000_000)
    ap.add_argument("--ref-procs", type=int, default=0, help="host processes for the reference arm (0 = all that fit)")
    ap.add_argument("--ref-sample", type=int, default=60000, help="active particles per reference process per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    if args.impl == "reference":
        main_reference(args)
    else:
        main_b200(args)


if __name__ == "__main__":
    main()
