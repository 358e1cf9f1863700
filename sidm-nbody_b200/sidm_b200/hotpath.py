"""Host-side mirror of the reference's hot-path entry points, over the C ABI.

Method names and argument meaning follow the reference (accel.c:27 compute_accelerations,
gravtree.c:18 gravity_tree, forcetree.c:90 force_treebuild, sidm.c:57 sidm, sidm.c:814
sidm_ensure_neighbours, init.c:431 setup_smoothinglengths_sidm, sidm.c:970 getvmax,
forcetree.c:2311 ngb_treefind), so the parity tests read like calls into the reference.
Particle indices are 0-based (the reference's P[i+1]).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import B200Error, Counters, Params, Replay, check, ptr

DEFAULT_PARAMS = dict(TreeAllocFactor=0.8, ErrTolTheta=0.5, ErrTolForceAcc=0.005, TypeOfOpeningCriterion=1,
                      ComovingIntegrationOn=0, G=43007.1, SofteningHalo=0.3, BoxSize=0.0, PeriodicBoundariesOn=0,
                      Omega0=1.0, OmegaLambda=0.0, Hubble=0.1, DesNumNgb=30, MaxNumNgbDeviation=2,
                      CrossSectionInternal=2.089, CrossSectionType=0, YukawaVelocity=0.0, CrossSectionPowLaw=0.0,
                      CrossSectionVelScale=1.0, Seed=55, BunchSizeSidm=0, ReferenceNgbOrder=0)


def make_params(max_part, device=0, **kw):
    cfg = dict(DEFAULT_PARAMS)
    cfg.update(kw)
    soft = cfg.pop("SofteningHalo")
    table = cfg.pop("SofteningTable", None)
    p = Params(device=device, MaxPart=int(max_part), **cfg)
    for t in range(6):
        p.SofteningTable[t] = 0.0
    p.SofteningTable[1] = soft
    if table is not None:
        for t in range(6):
            p.SofteningTable[t] = table[t]
    return p


def _f32(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, np.float32)
    if shape is not None:
        assert a.shape == shape, (a.shape, shape)
    return a


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, np.int32)


class HotPath:
    """One device context (one process = one GPU, like one MPI rank of the reference)."""

    def __init__(self, max_part, device=0, **params):
        self.lib = capi.load()
        self.params = make_params(max_part, device, **params)
        check(self.lib.b200_init(C.byref(self.params)), "b200_init")
        self.n = 0
        self.time = 0.0
        self._aos = None

    def close(self):
        if self.lib is not None:
            self.lib.b200_finalize()
            self.lib = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_params(self, **kw):
        for k, v in kw.items():
            if k == "SofteningHalo":
                self.params.SofteningTable[1] = v
            else:
                setattr(self.params, k, v)
        check(self.lib.b200_set_params(C.byref(self.params)), "b200_set_params")

    def set_option(self, name, value):
        check(self.lib.b200_set_option(name.encode(), int(value)), "b200_set_option")

    # ---- particle state ---------------------------------------------------------------
    def set_particles(self, pos=None, vel=None, mass=None, ids=None, curtime=None, accel=None, oldacc=None,
                      hsml=None, dvel=None, n=None):
        n = n if n is not None else (len(mass) if mass is not None else self.n)
        pos, vel, accel, dvel = _f32(pos), _f32(vel), _f32(accel), _f32(dvel)
        mass, curtime, oldacc, hsml, ids = _f32(mass), _f32(curtime), _f32(oldacc), _f32(hsml), _i32(ids)
        check(self.lib.b200_set_soa(n, ptr(pos), ptr(vel), ptr(mass), ptr(ids), ptr(curtime), ptr(accel),
                                    ptr(oldacc), ptr(hsml), ptr(dvel)), "b200_set_soa")
        self.n = n

    def get(self, *names):
        """names from: PosPred VelPred Accel OldAcc GravCost HsmlVelDisp NgbVelDisp dVel Left Right"""
        n = self.n
        spec = dict(PosPred=((n, 3), np.float32), VelPred=((n, 3), np.float32), Accel=((n, 3), np.float32),
                    OldAcc=((n,), np.float32), GravCost=((n,), np.float32), HsmlVelDisp=((n,), np.float32),
                    NgbVelDisp=((n,), np.int32), dVel=((n, 3), np.float32), Left=((n,), np.float32),
                    Right=((n,), np.float32))
        order = ["PosPred", "VelPred", "Accel", "OldAcc", "GravCost", "HsmlVelDisp", "NgbVelDisp", "dVel", "Left", "Right"]
        out = {k: np.empty(*spec[k]) for k in names}
        check(self.lib.b200_get_soa(*[ptr(out.get(k)) for k in order]), "b200_get_soa")
        return out[names[0]] if len(names) == 1 else tuple(out[k] for k in names)

    def bind_particles(self, aos, pin=True):
        """aos: numpy structured array with capi.PARTICLE_DTYPE (= the reference's &P[1])."""
        assert aos.flags["C_CONTIGUOUS"]
        self._aos = aos
        lay = capi.layout_of(aos.dtype)
        check(self.lib.b200_bind_particles(aos.ctypes.data_as(C.c_void_p), len(aos), C.byref(lay), int(pin)),
              "b200_bind_particles")
        self.n = len(aos)

    def upload(self):
        check(self.lib.b200_upload(), "b200_upload")

    def download(self, into=None):
        if into is None:
            check(self.lib.b200_download(), "b200_download")
        else:
            assert into.dtype == self._aos.dtype and len(into) >= self.n and into.flags["C_CONTIGUOUS"]
            check(self.lib.b200_download_to(into.ctypes.data_as(C.c_void_p)), "b200_download_to")

    def upload_active(self, idx):
        """host -> device of what the driver changes between force computations (Pos Vel CurrentTime MaxPredTime; VelPred = Vel,
        dVel = 0) for the listed particles only (gravtree.c:149-166 moves 20 bytes per active particle)"""
        a = _i32(idx)
        check(self.lib.b200_upload_active(ptr(a), len(a)), "b200_upload_active")

    def download_active(self, idx, into=None):
        """device -> host of the fields the path writes for the listed particles and the partners kicked since the last download"""
        a = _i32(idx)
        dst = None
        if into is not None:
            assert into.dtype == self._aos.dtype and len(into) >= self.n and into.flags["C_CONTIGUOUS"]
            dst = into.ctypes.data_as(C.c_void_p)
        check(self.lib.b200_download_active(ptr(a), len(a), dst), "b200_download_active")

    def advance(self, active=None, time=None, count=False):
        """advance(), predict.c:245: leap-frog kick+drift of the active particles, clears dVel"""
        t = self.time if time is None else float(time)
        a = _i32(active)
        ns = C.c_int(0)
        check(self.lib.b200_advance(ptr(a), 0 if a is None else len(a), t, C.byref(ns) if count else None), "b200_advance")
        return ns.value

    def compute_potential(self):
        """compute_potential(), potential.c:18: P[].Potential of every particle (rebuilds the tree)"""
        out = np.empty(self.n, np.float32)
        check(self.lib.b200_compute_potential(ptr(out)), "b200_compute_potential")
        return out

    def compute_global_quantities_of_system(self):
        """global.c:18: SysState (energies, momenta, centre of mass per type and in total) from the device state;
        P[].Potential as the last compute_potential() left it"""
        from .capi import SysState
        st = SysState()
        check(self.lib.b200_compute_global_quantities(C.byref(st)), "b200_compute_global_quantities")
        return st

    def savepositions(self, path, time=None, mass_table=None, hubble_param=0.0):
        """savepositions(), io.c:16: GADGET format-1 snapshot file from the device state; returns header1.npart"""
        t = self.time if time is None else float(time)
        mt = None if mass_table is None else np.ascontiguousarray(mass_table, np.float64)
        assert mt is None or mt.size == 6
        npart = np.zeros(6, np.int32)
        check(self.lib.b200_savepositions(str(path).encode(), t, ptr(mt), float(hubble_param), ptr(npart)), "b200_savepositions")
        return npart

    def savepositions_part(self, path, first, count, num_files, time=None, mass_table=None, hubble_param=0.0):
        """one file of a snapshot split over NumFilesPerSnapshot files (io.c:90-103): rows [first, first+count) of the particle order"""
        t = self.time if time is None else float(time)
        mt = None if mass_table is None else np.ascontiguousarray(mass_table, np.float64)
        npart = np.zeros(6, np.int32)
        check(self.lib.b200_savepositions_part(str(path).encode(), t, ptr(mt), float(hubble_param), int(first), int(count), int(num_files), ptr(npart)),
              "b200_savepositions_part")
        return npart

    def rng_state(self):
        """the whole generator state of the path: (sidm() calls, find_timesteps() calls) - save it with a restart file"""
        st = np.zeros(2, np.uint64)
        check(self.lib.b200_get_rng_state(ptr(st)), "b200_get_rng_state")
        return st

    def set_rng_state(self, state):
        st = np.ascontiguousarray(state, np.uint64)
        assert st.size == 2
        check(self.lib.b200_set_rng_state(ptr(st)), "b200_set_rng_state")

    def read_ic(self, path):
        """read_ic() + init() start-up state (read_ic.c:32, init.c:76-100) from one format-1 file into the device state;
        returns (time, mass_table, npart)"""
        t = C.c_double(0.0)
        mt = np.zeros(6, np.float64)
        npart = np.zeros(6, np.int32)
        check(self.lib.b200_load_snapshot(str(path).encode(), C.byref(t), ptr(mt), ptr(npart)), "b200_load_snapshot")
        self.n = int(npart.sum())
        self.time = t.value
        return t.value, mt, npart

    def force_treeevaluate_potential(self, targets):
        t = _i32(targets)
        out = np.empty(len(t), np.float64)
        check(self.lib.b200_potential_raw(ptr(t), len(t), ptr(out)), "b200_potential_raw")
        return out

    def reflect(self, radius, active=None):
        """reflect(), reflection.c:7: returns the number of reflected particles"""
        a = _i32(active)
        n = C.c_int(0)
        check(self.lib.b200_reflect(ptr(a), 0 if a is None else len(a), float(radius), C.byref(n)), "b200_reflect")
        return n.value

    def set_field(self, name, arr):
        arr = np.ascontiguousarray(arr)
        check(self.lib.b200_set_field(name.encode(), arr.ctypes.data_as(C.c_void_p), arr.nbytes), "b200_set_field")

    def find_timesteps(self, mode=0, active=None, time=None, vmax=0.0, crit=0, eta=0.05, velscale=10.0, probtol=0.2, dyntol=0.05,
                       dtmax=1e30, dtmin=0.0, jitter=None):
        """find_timesteps(mode), timestep.c:17: new MaxPredTime of the active particles; returns (MaxPredTime per active entry, clamped)"""
        t = self.time if time is None else float(time)
        a = _i32(active)
        na = self.n if a is None else len(a)
        tp = capi.TimestepParams(int(crit), eta, velscale, probtol, dyntol, dtmax, dtmin)
        out = np.empty(na, np.float32)
        nc = C.c_int(0)
        jit = None if jitter is None else np.ascontiguousarray(jitter, np.float64)
        check(self.lib.b200_find_timesteps(ptr(a), 0 if a is None else len(a), int(mode), t, float(vmax), C.byref(tp), ptr(jit),
                                           ptr(out), C.byref(nc)), "b200_find_timesteps")
        return out, nc.value

    # ---- the hot path -----------------------------------------------------------------
    def predict_collisionless_only(self, time):
        self.time = float(time)
        check(self.lib.b200_predict(float(time)), "b200_predict")

    def force_treebuild(self):
        check(self.lib.b200_tree_build(), "b200_tree_build")
        return self.counters().num_nodes

    def gravity_tree(self, active=None, time=None):
        """walk + epilogue for the active list (None = every particle); the tree must be current."""
        t = self.time if time is None else float(time)
        a = _i32(active)
        check(self.lib.b200_gravity(ptr(a), 0 if a is None else len(a), t), "b200_gravity")

    def sidm(self, active=None, time=None, vmax=0.0, replay_rand=None, replay_dir=None, replay_extra=None, replay_extra_off=None):
        t = self.time if time is None else float(time)
        a = _i32(active)
        rp = None
        if replay_rand is not None:
            self._rr = np.ascontiguousarray(replay_rand, np.float64)
            self._rd = np.ascontiguousarray(replay_dir, np.float64)
            rp = Replay(rand=self._rr.ctypes.data, dir=self._rd.ctypes.data)
            if replay_extra_off is not None:
                self._rx = np.ascontiguousarray(replay_extra if len(replay_extra) else np.zeros(2), np.float64)
                self._ro = np.ascontiguousarray(replay_extra_off, np.int32)
                rp.extra, rp.extra_off = self._rx.ctypes.data, self._ro.ctypes.data
        check(self.lib.b200_sidm(ptr(a), 0 if a is None else len(a), t, float(vmax),
                                 C.byref(rp) if rp is not None else None), "b200_sidm")

    def setup_nbr_sidm(self, active=None):
        a = _i32(active)
        check(self.lib.b200_setup_nbr_sidm(ptr(a), 0 if a is None else len(a)), "b200_setup_nbr_sidm")

    def sidm_ensure_neighbours(self, mode=0, time=None, vmax=0.0):
        t = self.time if time is None else float(time)
        check(self.lib.b200_sidm_ensure_neighbours(int(mode), t, float(vmax), None), "b200_sidm_ensure_neighbours")

    def setup_smoothinglengths_sidm(self, desired_ngb=30):
        check(self.lib.b200_setup_smoothinglengths_sidm(int(desired_ngb)), "b200_setup_smoothinglengths_sidm")

    def compute_accelerations(self, mode=0, active=None, time=None, vmax=0.0):
        t = self.time if time is None else float(time)
        self.time = t
        a = _i32(active)
        check(self.lib.b200_compute_accelerations(int(mode), ptr(a), 0 if a is None else len(a), t, float(vmax)),
              "b200_compute_accelerations")

    def getvmax(self):
        v = C.c_double(0)
        check(self.lib.b200_getvmax(C.byref(v)), "b200_getvmax")
        return v.value

    def ngb_treefind(self, idx, desngb=30):
        idx = _i32(idx)
        out = np.empty(len(idx), np.float32)
        check(self.lib.b200_ngb_treefind(ptr(idx), len(idx), int(desngb), ptr(out)), "b200_ngb_treefind")
        return out

    # ---- parity / debug ---------------------------------------------------------------
    def force_treeevaluate_direct(self, targets):
        t = _i32(targets)
        acc = np.empty((len(t), 3), np.float64)
        check(self.lib.b200_direct(ptr(t), len(t), ptr(acc)), "b200_direct")
        return acc

    def force_treeevaluate(self, targets):
        """raw double accelerations (pre-G) and (particle, node) interaction counts per target"""
        t = _i32(targets)
        acc = np.empty((len(t), 3), np.float64)
        cost = np.empty((len(t), 2), np.int32)
        check(self.lib.b200_walk_raw(ptr(t), len(t), ptr(acc), ptr(cost)), "b200_walk_raw")
        return acc, cost

    def get_tree(self):
        m = C.c_int(0)
        check(self.lib.b200_get_tree(C.byref(m), *([None] * 9)), "b200_get_tree")
        m = m.value
        d = dict(center=np.empty((m, 3), np.float32), len=np.empty(m, np.float32), mass=np.empty(m, np.float32),
                 s=np.empty((m, 3), np.float32), Q=np.empty((m, 7), np.float32), oc=np.empty(m, np.float32),
                 bmax2=np.empty(m, np.float32), count=np.empty(m, np.int32), level=np.empty(m, np.int32))
        mm = C.c_int(0)
        check(self.lib.b200_get_tree(C.byref(mm), ptr(d["center"]), ptr(d["len"]), ptr(d["mass"]), ptr(d["s"]),
                                     ptr(d["Q"]), ptr(d["oc"]), ptr(d["bmax2"]), ptr(d["count"]), ptr(d["level"])),
              "b200_get_tree")
        return d

    def ngb_lists(self, idx, cap=512):
        idx = _i32(idx)
        cnt = np.empty(len(idx), np.int32)
        lst = np.full((len(idx), cap), -1, np.int32)
        check(self.lib.b200_ngb_lists(ptr(idx), len(idx), cap, ptr(cnt), ptr(lst)), "b200_ngb_lists")
        return cnt, lst

    def sidm_debug(self, nslot):
        sp = np.empty(nslot, np.int32)
        pmax = np.empty(nslot, np.float64)
        prob = np.empty(nslot, np.float64)
        partner = np.empty(nslot, np.int32)
        check(self.lib.b200_sidm_debug(nslot, ptr(sp), ptr(pmax), ptr(prob), ptr(partner)), "b200_sidm_debug")
        return sp, pmax, prob, partner

    def scatlog(self, cap=1 << 20):
        out = np.zeros(cap, capi.SCATLOG_DTYPE)
        n = C.c_int(0)
        check(self.lib.b200_get_scatlog(ptr(out), cap, C.byref(n)), "b200_get_scatlog")
        return out[:n.value].copy()

    def counters(self):
        c = Counters()
        check(self.lib.b200_get_counters(C.byref(c)), "b200_get_counters")
        return c

    def peek(self, name, dtype, shape):
        """copy an internal device buffer (b200_device_buffer name) to the host - debugging / bench set-up"""
        addr, nb = self.device_buffer(name)
        out = np.empty(shape, dtype)
        assert out.nbytes <= nb, (out.nbytes, nb)
        rt = C.CDLL("libcudart.so.12")
        rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        rc = rt.cudaMemcpy(out.ctypes.data_as(C.c_void_p), C.c_void_p(addr), out.nbytes, 2)
        if rc != 0:
            raise B200Error(9001, f"cudaMemcpy({name}) -> {rc}")
        return out

    def device_buffer(self, name):
        p = C.c_void_p(0)
        nb = C.c_longlong(0)
        check(self.lib.b200_device_buffer(name.encode(), C.byref(p), C.byref(nb)), "b200_device_buffer")
        return p.value, nb.value
