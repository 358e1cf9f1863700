#!/usr/bin/env python
"""bench.py - particle-updates/s of the tree-gravity + SIDM hot path (BASELINE.json metric).

Default workload = BASELINE.json configs[2] ("C3", the one the metric is quoted on; it fits one B200):
isolated NFW halo of parameter.txt:6-11, N = 1e7 equal-mass particles, sigma/m = 1 cm^2/g,
eps = 0.3 kpc, ErrTolTheta 0.5, relative opening criterion alpha = 0.005, DesNumNgb 30 +- 2,
every particle active on a fixed step (all-active steps).  Synthetic seeded
ICs (sidm_b200/ic.py), the reference's own parameter values.  `--config C1|C2|C3|C3S|C4|C5`
selects the other BASELINE configurations (same JSON line):
  C1  Hernquist N = 1e5, sigma/m = 1                         (configs[0], the reference's CPU-runnable case)
  C2  NFW N = 1e6, sigma/m = 1                               (configs[1]; tests/test_gpu_fullsize.py holds its direct-sum check)
  C3S C3 at sigma/m = 10 and a long step: >= 1e4 scatterings per step, so the pair pass (k_pass2) and the write sweeps are timed
  C4  periodic comoving box 256^3, Ewald correction, sigma/m = 0.5 cm^2/g (h^-1 Mpc units)   (configs[3])
  C5  NFW N = 4e6 three times as concentrated, sigma/m = 10, individual time steps: the active lists of the reference's own
      time line captured at N = 32768 (tests/golden/make_c5_timeline.py) and scaled by radius rank   (configs[4])

A "step" = compute_accelerations(0) (predict + tree build + walk + sidm + ensure_neighbours,
accel.c:27-132) for the active particles followed by advance() (predict.c:245) so that the next
step sees moved particles.  value = active particles * K / time with the particle state resident in HBM;
e2e = the same through the drop-in boundary: host array-of-structs (124-byte particle_data)
-> b200_upload -> b200_compute_accelerations -> b200_download (all-active configs: whole array; C5: the active
particles only, b200_upload_active / b200_download_active), copies inside the timed region.

`--impl reference` times the UNMODIFIED reference (oracle/_ref/libsidmref_fast.so or, for C4, the -DPERIODIC build,
compiled from /root/reference by oracle/Makefile) on the host cores: P forked single-rank copies each hold
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
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "sidm-nbody_b200"))

METRIC = "particle-updates/s (tree gravity+SIDM DSMC), N=1e7 halo, 1/2/4/8 B200"
UNIT = "particle-updates/s"
FP32_PEAK_TFLOPS = 74.0    # FFMA rate measured on this pool's B200 by scripts/ubench/ffma2.cu (profiles/ffma2_ubench.txt)

# box of C4: total mass such that Omega0 = 0.3 (check_omega(), init.c:201-225): rho_crit = 3 H^2 / (8 pi G)
_G, _H, _O0, _BOX = 43007.1, 0.1, 0.3, 100.0
C5_IC = dict(seed=5, rho0=1.49e-4 * 27, rs=11.14356 / 3)

CONFIGS = {
    "C1": dict(n=100_000, ic="hernquist", ic_kw=dict(seed=1), sigma=1.0, dt=1.0e-3,
               workload="Hernquist halo N={n:.0e} (M=1e10 Msun, a=10 kpc, rmax=100 a, seed 1), sigma/m=1 cm^2/g, eps=0.3 kpc, all-active steps dt=1e-3"),
    "C2": dict(n=1_000_000, ic="nfw", ic_kw=dict(seed=2), sigma=1.0, dt=3.0e-4,
               workload="NFW halo N={n:.0e} (rho0=1.49e-4, rs=11.14 kpc, rmax=100 rs, seed 2), sigma/m=1 cm^2/g, eps=0.3 kpc, all-active steps dt=3e-4"),
    # dt of C3: the G*rho / SIDM step limit of timestep.c:247-265 at the centre of this halo at N=1e7 (internal time unit 0.978 Gyr),
    # so that an all-active step is one the reference would take
    "C3": dict(n=10_000_000, ic="nfw", ic_kw=dict(seed=3), sigma=1.0, dt=1.0e-4,
               workload="NFW halo N={n:.0e} (rho0=1.49e-4, rs=11.14 kpc, rmax=100 rs, seed 3), sigma/m=1 cm^2/g, eps=0.3 kpc, all-active steps dt=1e-4"),
    "C3S": dict(n=10_000_000, ic="nfw", ic_kw=dict(seed=3), sigma=10.0, dt=5.0e-2,
                workload="NFW halo N={n:.0e} (seed 3), sigma/m=10 cm^2/g, eps=0.3 kpc, all-active steps dt=5e-2 (scatter-heavy: >=1e4 scatterings per step)"),
    "C4": dict(n=256 ** 3, ic="periodic", ic_kw=dict(seed=4, box=_BOX, vel_sigma=30.0), sigma=0.5, dt=1.0e-4, periodic=True, ref_kind="periodic",
               workload="periodic comoving box {ng}^3 = {n} particles (L=100 h^-1 Mpc, Omega0=0.3, OmegaLambda=0.7, a=0.1, grid + 0.2-cell Gaussian displacements, seed 4), Ewald correction, sigma/m=0.5 cm^2/g, all-active steps da=1e-4"),
    "C5": dict(n=4_000_000, ic="nfw", ic_kw=C5_IC, sigma=10.0, dt=None, timeline=True,
               workload="NFW halo N={n:.0e} three times as concentrated as parameter.txt's (seed 5), sigma/m=10 cm^2/g, individual time steps: active lists of the reference's own time line (captured at N=32768, scaled by radius rank)"),
}


def make_ic(cfg, n):
    from sidm_b200 import ic
    if cfg["ic"] == "hernquist":
        return ic.hernquist(n, **cfg["ic_kw"])
    if cfg["ic"] == "nfw":
        return ic.nfw(n, **cfg["ic_kw"])
    ng = int(round(n ** (1.0 / 3)))
    mtot = _O0 * 3 * _H * _H / (8 * np.pi * _G) * _BOX ** 3
    return ic.periodic_box(ng, total_mass=mtot, **cfg["ic_kw"])


def sigma_internal(cfg):
    from sidm_b200 import ic
    return ic.cross_section_internal(cfg["sigma"], unit_length_cm=3.085678e24 if cfg.get("periodic") else 3.085678e21)


def path_params(cfg, n):
    """b200_params / reference parameters that differ from the sample parameter file"""
    if not cfg.get("periodic"):
        return {}
    ng = int(round(n ** (1.0 / 3)))
    return dict(BoxSize=_BOX, PeriodicBoundariesOn=1, SofteningHalo=_BOX / ng / 25, ComovingIntegrationOn=1, Omega0=_O0, OmegaLambda=0.7, Hubble=_H)


def t_begin(cfg):
    return 0.1 if cfg.get("periodic") else 0.0


class Timeline:
    """the reference's captured time line of C5 (tests/golden/c5_timeline.npz), scaled to n particles by radius rank"""

    def __init__(self, n, pos):
        d = np.load(os.path.join(ROOT, "tests", "golden", "c5_timeline.npz"))
        self.nred = int(d["nred"]); self.time = d["time"]; self.off = d["offsets"]; self.ranks = d["ranks"]
        self.n = n
        self.order = np.argsort((pos.astype(np.float64) ** 2).sum(1), kind="stable").astype(np.int32)   # radius rank -> particle
        self.iters = len(self.time)

    def active(self, it):
        it = it % self.iters
        r = self.ranks[self.off[it]:self.off[it + 1]].astype(np.int64)
        f = self.n // self.nred
        idx = (r[:, None] * f + np.arange(f)[None, :]).ravel()
        return np.sort(self.order[idx]).astype(np.int32)

    def t(self, it):
        # the captured times repeat with a constant offset when the list wraps around
        span = self.time[-1] - self.time[0] + (self.time[-1] - self.time[-2])
        return float(self.time[it % self.iters] + (it // self.iters) * span)


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

def _ref_worker(conn, cfg, n, pos, vel, mass, ids, sample, pool_desc):
    """one single-rank copy of the unmodified reference; advances `sample` through compute_accelerations(0)"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refdrv
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)                                  # the reference prints progress lines to stdout
    try:
        import tempfile
        os.chdir(tempfile.mkdtemp())                     # the periodic build caches its Ewald table in the working directory
        R = refdrv.Reference(cfg.get("ref_kind", "fast"))
        kw = dict(path_params(cfg, n))
        kw.pop("PeriodicBoundariesOn", None)
        t0 = t_begin(cfg)
        R.setup(n, CrossSectionInternal=sigma_internal(cfg), TreeUpdateFrequency=0.1, BufferSizeMB=100, Time=t0, **kw)
        R.init_rand(55)
        R.set_particles(pos, vel, mass, ids)
        R.treebuild()
        h = np.zeros(n, np.float32)
        for i in sample:
            h[i] = np.sqrt(R.ngb_treefind(pos[i], 30))
        R.set("HSML", h)
        R.set("NGB", np.full(n, 30, np.int32))           # only the sample is ever out of range
        dt = cfg["dt"] if cfg["dt"] else 2.4e-5          # C5: the median step of the captured time line
        R.all_active(t0, t0 + dt / 2)                    # CurrentTime t0, prediction time t0 + dt/2
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


def run_reference_sample(cfg, n, nproc, per_proc, warmup, steps, ic_data=None):
    """returns (updates_per_s, cores, sample_description, seconds_per_step)"""
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refdrv
    if not refdrv.available(cfg.get("ref_kind", "fast")):
        return None
    pos, vel, mass, ids = ic_data if ic_data is not None else make_ic(cfg, n)
    rng = np.random.default_rng(12345)
    pool, pool_desc = np.arange(n), "all particles"
    if cfg.get("timeline"):                               # C5: sample from one of the captured active lists
        tl = Timeline(n, pos)
        pool = tl.active(int(np.argmax(np.diff(tl.off))))
        pool_desc = f"the largest captured active list ({len(pool)} particles)"
    allsample = rng.choice(pool, size=min(len(pool), nproc * per_proc), replace=False).astype(np.int32)
    parts = np.array_split(allsample, nproc)
    ctx = mp.get_context("fork")
    procs = []
    for p in range(nproc):
        a, b = ctx.Pipe()
        pr = ctx.Process(target=_ref_worker, args=(b, cfg, n, pos, vel, mass, ids, np.sort(parts[p]), pool_desc), daemon=True)
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
    desc = (f"{len(allsample)} of {n} particles active per step, drawn from {pool_desc} ({nproc} forked single-rank copies of the unmodified reference x "
            f"{len(parts[0])} targets, each holding the full {n}-particle tree; shipped TreeUpdateFrequency 0.1 so the tree build is amortised as in the reference); "
            "a sampled estimate: the per-call O(N) work (prediction of all N, node updates) is spread over the sample only, which favours neither side at this ratio")
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


def main_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.n
    nproc = args.ref_procs or host_parallelism(n)
    t0 = time.time()
    r = run_reference_sample(cfg, n, nproc, args.ref_sample, args.warmup, args.steps)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref reference library not built (needs /root/reference at build time)"}))
        return
    value, cores, desc, tstep = r
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": tstep * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(args, cfg), "config": args.config, "particles": n, "parallelism": f"host cpu x{cores}"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.time() - t0}
    print(json.dumps(line))


def workload_name(args, cfg):
    return cfg["workload"].format(n=args.n, ng=int(round(args.n ** (1.0 / 3))))


# ------------------------------------------------------------------------------------ B200 arm

ROOF_STEPS = 2


def state_crc(hp, n):
    """CRC-32 of the particle state the path owns (bits of Pos, Vel, Accel, OldAcc, HsmlVelDisp, NgbVelDisp after the timed steps):
    the same number at every GPU count proves that the sharded run is bit-identical to the one-GPU run"""
    crc = 0
    for name, dt, shape in (("pos0", np.float32, (n, 3)), ("velh", np.float32, (n, 4)), ("accel", np.float32, (n, 3)), ("oldacc", np.float32, (n,)),
                            ("ngb", np.int32, (n,)), ("curtime", np.float32, (n,))):
        crc = zlib.crc32(hp.peek(name, dt, shape).tobytes(), crc)
    return f"{crc & 0xffffffff:08x}"


def main_b200(args, cfg):
    import torch
    import torch.distributed as dist
    from sidm_b200 import HotPath

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 arm has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.n
    pos, vel, mass, ids = make_ic(cfg, n)
    n = len(mass)
    hp = HotPath(n, device=local, CrossSectionInternal=sigma_internal(cfg), Seed=55, **path_params(cfg, n))

    from sidm_b200.multi import Sharder
    sh = Sharder(hp, world, rank)

    # ---- set-up (untimed): start-up forces and smoothing lengths, as init.c:120-180
    T0 = t_begin(cfg)
    hp.set_particles(pos, vel, mass, ids, curtime=np.full(n, T0, np.float32))
    hp.predict_collisionless_only(T0)
    hp.force_treebuild()
    hp.setup_smoothinglengths_sidm(30)
    vmax = hp.getvmax()
    sh.compute_accelerations(1, time=T0, vmax=vmax)       # BH criterion (OldAcc = 0) -> OldAcc

    tl = Timeline(n, pos) if cfg.get("timeline") else None
    reuse = max(args.tree_reuse, 0)
    hp.set_option("tree_reuse", reuse)
    state = dict(t=T0, it=0, active=0)

    def step():
        if tl is None:
            t = state["t"] + cfg["dt"] / 2
            sh.compute_accelerations(0, time=t, vmax=vmax)
            hp.advance(time=t)
            state["t"] += cfg["dt"]; state["active"] += n
        else:
            act = tl.active(state["it"]); t = tl.t(state["it"])
            sh.compute_accelerations(0, time=t, vmax=vmax, active=act)
            hp.advance(active=act, time=t)
            state["it"] += 1; state["active"] += len(act)

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
    state["active"] = 0
    # inputs smaller than L2 (C1): every step is timed on its own and 256 MB are written in between (not timed)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if n < 1_000_000 else None
    ms_steps = 0.0
    ev0.record()
    for _ in range(args.steps):
        if flush is not None:
            flush.fill_(1); torch.cuda.synchronize()
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
        step()
        if flush is not None:
            eb.record(); torch.cuda.synchronize(); ms_steps += ea.elapsed_time(eb)
        c = hp.counters()
        walk_ms.append(c.ms_walk); build_ms.append(c.ms_build); sidm_ms.append(c.ms_sidm); ens_ms.append(c.ms_ensure)
        list_nodes += c.list_nodes; list_parts += c.list_parts; ntarg += c.num_targets; num_lists += c.num_lists
        inter_p += c.part_interactions; inter_n += c.node_interactions
        scat += c.sct_scattered; rep_it += c.ensure_iterations; rep_n += c.ensure_repaired; cand += c.ngb_candidates
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1) if flush is None else ms_steps
    if world > 1:
        tt = torch.tensor([ms], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    clk = clocks.stop() if rank == 0 else None
    c1 = hp.counters()
    launches = (c1.kernel_launches - c0.kernel_launches)
    updates = state["active"]
    value = updates / (ms * 1e-3)
    crc = state_crc(hp, n)

    # ---- roofline of the dominant kernel (the tree walk), SURVEY.md 8d bytes formula.  In the timed
    # steps the walk shares the GPU with the SIDM chain (two streams), so its event-bracketed duration
    # there is not the kernel's own; the kernel duration is taken live from ROOF_STEPS further steps of
    # the same run with the two phases in sequence (b200_set_option("overlap", 0)).
    peak, peak_src = peaks()
    wms_overlapped = float(np.mean(walk_ms))
    hp.set_option("overlap", 0)
    iso = []
    r_nodes = r_parts = r_targ = r_lists = r_in = r_ip = 0
    for _ in range(ROOF_STEPS):
        step()
        c = hp.counters()
        iso.append(c.ms_walk)
        r_nodes += c.list_nodes; r_parts += c.list_parts; r_targ += c.num_targets; r_lists += c.num_lists
        r_in += c.node_interactions; r_ip += c.part_interactions
    hp.set_option("overlap", -1)                          # back to the default mode
    wms = float(np.mean(iso))
    a_per_launch = r_targ / ROOF_STEPS
    lists = max(1.0, r_lists / ROOF_STEPS)                # interaction lists per launch (one per warp of 32 targets)
    i_n = r_nodes / ROOF_STEPS / lists
    i_p = r_parts / ROOF_STEPS / lists
    walk_bytes = a_per_launch * 32 + lists * (48 * i_n + 16 * i_p)
    achieved = walk_bytes / (wms * 1e-3) / 1e9
    flops = (r_in * 70.0 + r_ip * 20.0) / ROOF_STEPS      # ~70 flop per cell interaction, ~20 per particle interaction
    traffic = None
    tj = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tj) and world == 1 and args.config == "C3":
        t = json.load(open(tj))
        if int(t.get("particles", 0)) == n:
            traffic = t["traffic_bytes_per_launch"]      # ncu dram read+write bytes of one k_walk launch (profiles/)
    tfl = flops / (wms * 1e-3) / 1e12
    roofline = {"bound": "hbm", "kernel": "k_walk", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "ms_per_launch": wms,
                "algorithmic_bytes_per_launch": walk_bytes, "I_n_per_list": i_n, "I_p_per_list": i_p, "targets_per_list": a_per_launch / lists,
                "interactions_per_target": {"node": r_in / max(1, r_targ), "particle": r_ip / max(1, r_targ)},
                "fp32": {"achieved_tflops": tfl, "peak_tflops": FP32_PEAK_TFLOPS, "frac": tfl / FP32_PEAK_TFLOPS,
                         "peak_source": "FFMA rate of scripts/ubench/ffma2.cu on this pool's B200 (profiles/ffma2_ubench.txt)",
                         "flop_model": "70 per cell interaction + 20 per particle interaction (SURVEY.md 8d)"},
                "ms_per_launch_while_sidm_overlaps": wms_overlapped,
                "note": "the walk is bound by instruction issue (FP32 + integer pipes), not by HBM: DESIGN.md section 5; both fractions are reported"}
    phases = {"build_ms": float(np.mean(build_ms)), "walk_ms": wms, "sidm_ms": float(np.mean(sidm_ms)), "ensure_ms": float(np.mean(ens_ms)),
              "scatterings_per_step": scat / args.steps, "ensure_passes_per_step": rep_it / args.steps,
              "ensure_repaired_per_step": rep_n / args.steps, "active_per_step": updates / args.steps,
              "ngb_candidates_per_search": cand / max(1.0, updates + rep_n)}

    # ---- end to end through the drop-in boundary (host AoS in pinned memory)
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(hp, sh, n, mass, ids, vmax, state, args, world, rank, step, cfg, tl)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            nproc = args.ref_procs or host_parallelism(n)
            r = run_reference_sample(cfg, n, nproc, max(500, args.ref_sample // 4), 1, 1, (pos, vel, mass, ids))
            if r is not None:
                cpu = {"value": r[0], "unit": UNIT, "cores": r[1], "kind": "reference", "sample": r[2]}
        except Exception as e:  # pragma: no cover
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"failed: {e!r}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": workload_name(args, cfg), "config": args.config, "particles": n, "parallelism": sh.describe(),
                           "l2": "inputs larger than L2 (particle + node records per step vs 126 MB L2)" if n >= 1_000_000 else
                                 "inputs smaller than L2: 256 MB written to HBM between the timed steps (each step timed on its own with CUDA events)",
                           "step": "compute_accelerations(0) + advance()" + (", all particles active" if tl is None else ", the time line's active particles"),
                           "tree_reuse": reuse},
                "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "phases": phases, "state_crc": crc}
        print(json.dumps(line), flush=True)
    if args.overlap_sweep:
        # A/B of the option "overlap" (0 = phases in sequence, 1 = SIDM chain next to the walk, 2 = SIDM pass next to the walk)
        # on the state the run has reached, after the line above: 2 untimed + 6 timed steps per mode, stderr only
        for m in (0, 1, 2, 1, 0):
            hp.set_option("overlap", m)
            for _ in range(2):
                step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(6):
                step()
            barrier()
            if rank == 0:
                print(f"bench.py overlap sweep: mode {m}: {(time.perf_counter() - t0) / 6 * 1e3:.3f} ms per step on {world} GPU(s)", file=sys.stderr, flush=True)
        hp.set_option("overlap", -1)
    hp.close()
    if world > 1:
        dist.destroy_process_group()


def run_e2e(hp, sh, n, mass, ids, vmax, state, args, world, rank, step_fn, cfg, tl):
    """host AoS -> upload -> compute_accelerations(0) -> download, every step, copies timed.
    All-active configs: the inputs are successive states of the run (one fresh host array per step), captured untimed.
    C5: one host array holds the run's state; each step uploads what the driver changed for the previous active list
    (b200_upload_active) and downloads the new active particles' results (b200_download_active)."""
    import torch
    from sidm_b200 import capi
    rt = torch.cuda.cudart()

    def host_state():
        a = np.zeros(n, capi.PARTICLE_DTYPE)
        velh = hp.peek("velh", np.float32, (n, 4))
        a["Pos"] = hp.peek("pos0", np.float32, (n, 3)); a["PosPred"] = a["Pos"]
        a["Vel"] = velh[:, :3]; a["VelPred"] = a["Vel"]; a["HsmlVelDisp"] = velh[:, 3]
        a["Mass"] = mass; a["ID"] = ids; a["Type"] = 1
        a["CurrentTime"] = hp.peek("curtime", np.float32, (n,))
        a["Accel"], a["OldAcc"], a["NgbVelDisp"] = hp.get("Accel", "OldAcc", "NgbVelDisp")
        return a

    if tl is not None:
        if world > 1:
            return None
        host = host_state()
        rt.cudaHostRegister(host.ctypes.data, host.nbytes, 0)
        hp.bind_particles(host, pin=False)
        hp.upload()
        steps = max(2, min(args.steps, 8))
        prev = np.zeros(0, np.int32)
        up = down = nact = 0

        def one(k):
            nonlocal prev, up, down, nact
            act = tl.active(state["it"]); t = tl.t(state["it"])
            hp.upload_active(prev)
            hp.compute_accelerations(0, active=act, time=t, vmax=vmax)
            hp.download_active(act)
            up += len(prev) * 36; down += len(act) * 76; nact += len(act)
            # the host driver's part, untimed work of the reference (advance() on the host array = what upload_active sends next)
            hp.advance(active=act, time=t)
            host["Pos"][act] = hp.peek("pos0", np.float32, (n, 3))[act]
            host["Vel"][act] = hp.peek("velh", np.float32, (n, 4))[act, :3]
            host["CurrentTime"][act] = hp.peek("curtime", np.float32, (n,))[act]
            prev = act
            state["it"] += 1

        one(0)
        up = down = nact = 0
        torch.cuda.synchronize()
        tsum = t_up = t_cmp = t_down = 0.0
        for k in range(steps):
            act = tl.active(state["it"]); t = tl.t(state["it"])
            t0 = time.perf_counter()
            hp.upload_active(prev)
            t1 = time.perf_counter()
            hp.compute_accelerations(0, active=act, time=t, vmax=vmax)
            t2 = time.perf_counter()
            hp.download_active(act)
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            tsum += t3 - t0; t_up += t1 - t0; t_cmp += t2 - t1; t_down += t3 - t2
            up += len(prev) * 36; down += len(act) * 76; nact += len(act)
            hp.advance(active=act, time=t)
            host["Pos"][act] = hp.peek("pos0", np.float32, (n, 3))[act]
            host["Vel"][act] = hp.peek("velh", np.float32, (n, 4))[act, :3]
            host["CurrentTime"][act] = hp.peek("curtime", np.float32, (n,))[act]
            prev = act
            state["it"] += 1
        rt.cudaHostUnregister(host.ctypes.data)
        return {"value": nact / tsum, "unit": UNIT, "h2d_bytes_per_step": int(up / steps), "d2h_bytes_per_step": int(down / steps),
                "steps": steps, "ms_per_step": tsum / steps * 1e3,
                "ms_upload": t_up / steps * 1e3, "ms_compute": t_cmp / steps * 1e3, "ms_download": t_down / steps * 1e3,
                "api": "b200_upload_active(previous active list: Pos Vel CurrentTime MaxPredTime, 36 B each) + b200_compute_accelerations(0, active) + "
                       "b200_download_active(active list + kicked partners, 76 B each) on a pinned 124-byte particle_data array"}

    nsnap = max(2, min(args.steps, 3))
    snaps, times = [], []
    for s in range(nsnap):
        snaps.append(host_state()); times.append(state["t"] + cfg["dt"] / 2)
        step_fn()
    out = np.zeros(n, capi.PARTICLE_DTYPE)
    for a in snaps + [out]:
        rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)
    hp.bind_particles(snaps[0], pin=False)

    parts = [0.0, 0.0, 0.0]

    def one(k):
        c0 = time.perf_counter()
        hp.bind_particles(snaps[k % nsnap], pin=False)
        sh.upload()                                  # own rows over PCIe (+ NVLink all-gather when sharded)
        c1 = time.perf_counter()
        sh.compute_accelerations(0, time=times[k % nsnap], vmax=vmax)
        c2 = time.perf_counter()
        sh.download(into=out)
        c3 = time.perf_counter()                     # each of the three calls returns after its own stream synchronisation
        parts[0] += c1 - c0; parts[1] += c2 - c1; parts[2] += c3 - c2
        return int(hp.counters().sct_scattered)

    steps = max(2, min(args.steps, 5))
    one(0)
    torch.cuda.synchronize()
    parts[:] = [0.0, 0.0, 0.0]
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
            "ms_upload": parts[0] / steps * 1e3, "ms_compute": parts[1] / steps * 1e3, "ms_download": parts[2] / steps * 1e3,   # rank 0's
            "api": ("b200_bind_particles + b200_upload + b200_compute_accelerations(0) + b200_download_to on pinned 124-byte particle_data arrays (successive states of the run)"
                    if world == 1 else "b200_upload_shard (own rows over PCIe + NVLink all-gather) + b200_compute_accelerations(0) + b200_download_shard, pinned 124-byte particle_data rows per rank")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS), help="BASELINE.json configuration (default C3 = the one the metric is quoted on)")
    ap.add_argument("--n", "--particles", dest="n", type=lambda s: int(float(s)), default=0, help="particle number (default: the configuration's)")
    ap.add_argument("--ref-procs", type=int, default=0, help="host processes for the reference arm (0 = all that fit)")
    ap.add_argument("--ref-sample", type=int, default=60000, help="active particles per reference process per step")
    ap.add_argument("--tree-reuse", type=int, default=0, help="option tree_reuse of the library: full build every k-th step, refits in between (default 0 = build at "
                    "every step, the only setting whose forces are the reference's to 1e-4; see tests/test_gpu_reuse.py)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--overlap-sweep", action="store_true", help="after the bench line: time the three values of the library option `overlap` (stderr)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if not args.n:
        args.n = cfg["n"]
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    if args.impl == "reference":
        main_reference(args, cfg)
    else:
        main_b200(args, cfg)


if __name__ == "__main__":
    main()
