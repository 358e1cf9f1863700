"""ctypes wrapper of oracle/liboracle.so - this repo's CPU restatement of the reference's hot
path (oracle_tree.c, oracle_sidm.c).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg, never by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")


class OParams(C.Structure):
    _fields_ = [("theta", C.c_double), ("alpha", C.c_double), ("criterion", C.c_int), ("eps", C.c_double),
                ("G", C.c_double), ("des_ngb", C.c_int), ("max_dev", C.c_int), ("sigma", C.c_double),
                ("xs_type", C.c_int), ("vc", C.c_double), ("pl_n", C.c_double), ("pl_v0", C.c_double)]


class OSidmOut(C.Structure):
    _fields_ = [("slot_particle", C.POINTER(C.c_int)), ("rand", C.POINTER(C.c_double)), ("dir", C.POINTER(C.c_double)),
                ("pmax", C.POINTER(C.c_double)), ("prob", C.POINTER(C.c_double)), ("partner", C.POINTER(C.c_int)),
                ("ngb", C.POINTER(C.c_int)), ("nslot", C.c_int), ("sct", C.c_int * 4), ("nlog", C.c_int),
                ("log_i", C.POINTER(C.c_int)), ("log_j", C.POINTER(C.c_int)), ("log_dv", C.POINTER(C.c_float)),
                ("extra", C.POINTER(C.c_double)), ("extra_off", C.POINTER(C.c_int))]


_lib = None


def lib():
    global _lib
    if _lib is None:
        src = [os.path.join(HERE, f) for f in ("oracle_tree.c", "oracle_sidm.c", "oracle.h")]
        if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in src):
            subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
        L = C.CDLL(LIB)
        L.otree_build.restype = C.c_void_p
        L.otree_build.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double]
        L.otree_free.argtypes = [C.c_void_p]
        L.otree_num_nodes.argtypes = [C.c_void_p]
        L.otree_random_subnodes.argtypes = [C.c_void_p]
        L.otree_dump.argtypes = [C.c_void_p] * 9
        L.otree_chain.argtypes = [C.c_void_p, C.c_void_p]
        L.otree_domain.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.otree_force.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.otree_direct.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ograv_epilogue.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ongb_variable.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int]
        L.ongb_treefind.restype = C.c_float
        L.ongb_treefind.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orng_new.restype = C.c_void_p
        L.orng_new.argtypes = [C.c_ulong, C.c_int]
        L.orng_free.argtypes = [C.c_void_p]
        L.orng_uniform.restype = C.c_double
        L.orng_uniform.argtypes = [C.c_void_p]
        L.orng_count.restype = C.c_long
        L.orng_count.argtypes = [C.c_void_p]
        L.ogetvmax.restype = C.c_double
        L.ogetvmax.argtypes = [C.c_int, C.c_void_p]
        L.osidm_pass.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 7 + [C.c_double, C.c_void_p, C.c_void_p]
        L.osidm_out_free.argtypes = [C.c_void_p]
        L.osidm_ensure.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 8 + [C.c_double, C.c_void_p]
        _lib = L
    return _lib


def global_quantities(pospred, velpred, mass, potential, types=None):
    """compute_global_quantities_of_system(), global.c:18-135: SysState as a flat array of 102 doubles
    (field order of allvars.h:517-537)"""
    L = lib()
    pp = np.ascontiguousarray(pospred, np.float32); vp = np.ascontiguousarray(velpred, np.float32)
    m = np.ascontiguousarray(mass, np.float32); pot = np.ascontiguousarray(potential, np.float32)
    ty = None if types is None else np.ascontiguousarray(types, np.int32)
    out = np.zeros(102, np.float64)
    L.oglobal_quantities.argtypes = [C.c_int] + [C.c_void_p] * 6
    L.oglobal_quantities.restype = None
    L.oglobal_quantities(len(m), _p(pp), _p(vp), _p(m), _p(pot), None if ty is None else _p(ty), _p(out))
    return out


def snapshot_bytes(pospred, velpred, ids, mass, types=None, time=0.0, mass_table=None, box=0.0, omega0=0.0,
                   omega_lambda=0.0, hubble_param=0.0, comoving=False, periodic=False, npart_total=None, num_files=1):
    """savepositions_ioformat1(), io.c:54-590, one rank, one file, no gas: the bytes of the snapshot file.
    Header struct io_header_1 (allvars.h:727-746, fill bytes zero), then PosPred, VelPred, ID, and the masses of the types
    whose MassTable entry is 0 - particles in type order (0..4; type 5 is not written, io.c:265), particle order
    inside a type - every block between int32 byte counts (io.c:207-210,261-262,576-578)."""
    pp = np.ascontiguousarray(pospred, np.float32).copy(); vp = np.ascontiguousarray(velpred, np.float32)
    ids = np.ascontiguousarray(ids, np.int32); m = np.ascontiguousarray(mass, np.float32)
    ty = np.ones(len(m), np.int32) if types is None else np.ascontiguousarray(types, np.int32)
    mt = np.zeros(6) if mass_table is None else np.asarray(mass_table, np.float64)
    if periodic:                                                    # io.c:275-283, float += double, one wrap per trip
        b = np.float64(box)
        for _ in range(64):
            lo = pp < 0
            if not lo.any():
                break
            pp[lo] = (pp[lo].astype(np.float64) + b).astype(np.float32)
        for _ in range(64):
            hi = pp.astype(np.float64) > b
            if not hi.any():
                break
            pp[hi] = (pp[hi].astype(np.float64) - b).astype(np.float32)
    order = np.concatenate([np.nonzero(ty == t)[0] for t in range(5)])
    cnt = np.array([(ty == t).sum() for t in range(5)] + [0], np.int32)
    assert cnt[0] == 0, "gas blocks are not part of this path"
    hdr = np.zeros(1, np.dtype([("npart", "<i4", 6), ("mass", "<f8", 6), ("time", "<f8"), ("redshift", "<f8"),
                                ("flag_sfr", "<i4"), ("flag_feedback", "<i4"), ("npartTotal", "<i4", 6),
                                ("flag_cooling", "<i4"), ("num_files", "<i4"), ("BoxSize", "<f8"), ("Omega0", "<f8"),
                                ("OmegaLambda", "<f8"), ("HubbleParam", "<f8"), ("flag_multiphase", "<i4"),
                                ("flag_stellarage", "<i4"), ("flag_sfrhistogram", "<i4"), ("fill", "S84")]))
    assert hdr.itemsize == 256
    hdr["npart"] = cnt; hdr["npartTotal"] = cnt if npart_total is None else np.asarray(npart_total, np.int32); hdr["mass"] = mt; hdr["time"] = time
    hdr["redshift"] = (1.0 / time - 1) if comoving else 0.0
    # io.c:140-160: npartTotal / num_files of a snapshot split over several files
    hdr["num_files"] = num_files; hdr["BoxSize"] = box; hdr["Omega0"] = omega0; hdr["OmegaLambda"] = omega_lambda
    hdr["HubbleParam"] = hubble_param
    withmass = np.concatenate([np.nonzero(ty == t)[0] for t in range(5) if mt[t] == 0] + [np.zeros(0, np.int64)]).astype(np.int64)

    def rec(payload):
        if len(payload) == 0:
            return b""
        mark = np.array([len(payload)], np.uint32).astype(np.int32, casting="unsafe").tobytes() if len(payload) >= 2**31 \
            else np.array([len(payload)], np.int32).tobytes()
        return mark + payload + mark
    return (rec(hdr.tobytes()) + rec(pp[order].tobytes()) + rec(vp[order].tobytes()) + rec(ids[order].tobytes())
            + rec(m[withmass].tobytes()))


def read_snapshot(path):
    """format-1 reader for round trips (read_ic.c:32-481 block order): header fields, pos, vel, id, mass (None if absent)"""
    raw = open(path, "rb").read()
    o = 0

    def block():
        nonlocal o
        nb = int(np.frombuffer(raw, "<i4", 1, o)[0]); o += 4
        body = raw[o:o + nb]; o += nb
        assert int(np.frombuffer(raw, "<i4", 1, o)[0]) == nb; o += 4
        return body
    h = block()
    npart = np.frombuffer(h, "<i4", 6, 0); mt = np.frombuffer(h, "<f8", 6, 24); time = float(np.frombuffer(h, "<f8", 1, 72)[0])
    n = int(npart.sum())
    pos = np.frombuffer(block(), "<f4").reshape(n, 3); vel = np.frombuffer(block(), "<f4").reshape(n, 3)
    ids = np.frombuffer(block(), "<i4")
    nm = int(sum(npart[t] for t in range(6) if mt[t] == 0))
    mass = np.frombuffer(block(), "<f4") if nm > 0 else None
    assert o == len(raw)
    return dict(npart=npart.copy(), mass_table=mt.copy(), time=time, pos=pos, vel=vel, ids=ids, mass=mass)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Holds one particle set; mirrors the reference's call sequence on it."""

    def __init__(self, pos, vel, mass, theta=0.5, alpha=0.005, criterion=1, eps=0.3, G=43007.1, des_ngb=30,
                 max_dev=2, sigma=2.089, xs_type=0, vc=0.0, pl_n=0.0, pl_v0=1.0):
        self.L = lib()
        self.pos = np.ascontiguousarray(pos, np.float32)
        self.vel = np.ascontiguousarray(vel, np.float32)
        self.mass = np.ascontiguousarray(mass, np.float32)
        self.n = len(self.mass)
        self.par = OParams(theta, alpha, criterion, eps, G, des_ngb, max_dev, sigma, xs_type, vc, pl_n, pl_v0)
        self.tree = None
        self.hsml = np.zeros(self.n, np.float32)
        self.dvel = np.zeros((self.n, 3), np.float32)
        self.ngb = np.zeros(self.n, np.int32)
        self.left = np.zeros(self.n, np.float32)
        self.right = np.zeros(self.n, np.float32)
        self.rng = None

    def __del__(self):
        try:
            if self.tree:
                self.L.otree_free(self.tree)
            if self.rng:
                self.L.orng_free(self.rng)
        except Exception:
            pass

    # ---- tree
    def treebuild(self):
        if self.tree:
            self.L.otree_free(self.tree)
        self.tree = self.L.otree_build(self.n, _p(self.pos), _p(self.mass), self.par.eps)
        return self.L.otree_num_nodes(self.tree)

    def random_subnodes(self):
        return self.L.otree_random_subnodes(self.tree)

    def dump(self):
        m = self.L.otree_num_nodes(self.tree)
        d = dict(center=np.empty((m, 3), np.float32), len=np.empty(m, np.float32), mass=np.empty(m, np.float32),
                 s=np.empty((m, 3), np.float32), Q=np.empty((m, 7), np.float32), oc=np.empty(m, np.float32),
                 bmax2=np.empty(m, np.float32), count=np.empty(m, np.int32))
        self.L.otree_dump(self.tree, *[_p(d[k]) for k in ("center", "len", "mass", "s", "Q", "oc", "bmax2", "count")])
        return d

    def chain(self):
        o = np.empty(self.n, np.int32)
        self.L.otree_chain(self.tree, _p(o))
        return o

    # ---- forces
    def force_tree(self, idx, oldacc=None):
        idx = np.ascontiguousarray(idx, np.int32)
        oa = None if oldacc is None else np.ascontiguousarray(oldacc, np.float32)
        acc = np.empty((len(idx), 3))
        cost = np.empty((len(idx), 2), np.int32)
        self.L.otree_force(self.tree, C.byref(self.par), len(idx), _p(idx), _p(oa), _p(acc), _p(cost))
        return acc, cost

    def potential(self, idx, oldacc=None):
        """raw tree potentials (forcetree.c:1389) and the float P[].Potential of compute_potential() (potential.c:131-168)"""
        idx = np.ascontiguousarray(idx, np.int32)
        oa = None if oldacc is None else np.ascontiguousarray(oldacc, np.float32)
        pot = np.empty(len(idx))
        self.L.otree_potential.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        self.L.otree_potential(self.tree, C.byref(self.par), len(idx), _p(idx), _p(oa), _p(pot))
        out = np.empty(len(idx), np.float32)
        m = np.ascontiguousarray(self.mass[idx])
        self.L.opot_epilogue.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        self.L.opot_epilogue(C.byref(self.par), len(idx), _p(pot), _p(m), _p(out))
        return pot, out

    def force_direct(self, idx):
        idx = np.ascontiguousarray(idx, np.int32)
        acc = np.empty((len(idx), 3))
        self.L.otree_direct(self.tree, C.byref(self.par), len(idx), _p(idx), _p(acc))
        return acc

    def epilogue(self, acc):
        acc = np.ascontiguousarray(acc, np.float64)
        a = np.empty((len(acc), 3), np.float32)
        oa = np.zeros(len(acc), np.float32)
        self.L.ograv_epilogue(C.byref(self.par), len(acc), _p(acc), _p(a), _p(oa))
        return a, oa

    # ---- neighbours
    def ngb_variable(self, xyz, h, cap=8192):
        xyz = np.ascontiguousarray(xyz, np.float32)
        lst = np.empty(cap, np.int32)
        r2 = np.empty(cap, np.float32)
        n = self.L.ongb_variable(self.tree, _p(xyz), C.c_float(h), _p(lst), _p(r2), cap)
        return lst[:n].copy(), r2[:n].copy()

    def ngb_treefind(self, xyz, desngb=30):
        xyz = np.ascontiguousarray(xyz, np.float32)
        return float(self.L.ongb_treefind(self.tree, _p(xyz), int(desngb)))

    # ---- SIDM
    def init_rand(self, seed):
        if self.rng:
            self.L.orng_free(self.rng)
        self.rng = self.L.orng_new(int(seed), 1)

    def rng_count(self):
        return self.L.orng_count(self.rng)

    def getvmax(self):
        return self.L.ogetvmax(self.n, _p(self.vel))

    def sidm(self, active, dt, vmax):
        """one sidm() call; dt = per-particle 2(t - t_i) (float32 array or scalar)"""
        active = np.ascontiguousarray(active, np.int32)
        dta = np.ascontiguousarray(np.broadcast_to(np.float32(dt), (self.n,)), np.float32)
        out = OSidmOut()
        self.L.osidm_pass(self.tree, C.byref(self.par), len(active), _p(active), _p(self.vel), _p(self.mass),
                          _p(self.hsml), _p(dta), _p(self.dvel), _p(self.ngb), float(vmax), self.rng, C.byref(out))
        ns, nl = out.nslot, out.nlog
        res = dict(slot_particle=np.ctypeslib.as_array(out.slot_particle, (ns,)).copy(),
                   rand=np.ctypeslib.as_array(out.rand, (ns,)).copy(),
                   dir=np.ctypeslib.as_array(out.dir, (ns * 3,)).reshape(ns, 3).copy(),
                   pmax=np.ctypeslib.as_array(out.pmax, (ns,)).copy(),
                   prob=np.ctypeslib.as_array(out.prob, (ns,)).copy(),
                   partner=np.ctypeslib.as_array(out.partner, (ns,)).copy(),
                   ngb=np.ctypeslib.as_array(out.ngb, (ns,)).copy(), sct=list(out.sct),
                   log_i=np.ctypeslib.as_array(out.log_i, (max(nl, 1),))[:nl].copy(),
                   log_j=np.ctypeslib.as_array(out.log_j, (max(nl, 1),))[:nl].copy(),
                   log_dv=np.ctypeslib.as_array(out.log_dv, (max(nl, 1) * 3,)).reshape(-1, 3)[:nl].copy())
        off = np.ctypeslib.as_array(out.extra_off, (ns + 1,)).copy()
        nx = int(off[-1])
        res["extra_off"] = off.astype(np.int32)
        res["extra"] = np.ctypeslib.as_array(out.extra, (max(nx, 1) * 2,))[:2 * nx].copy()
        self.L.osidm_out_free(C.byref(out))
        return res

    def find_timesteps(self, active, mode, time, vmax, accel, curtime, maxpred, crit=0, eta=0.05, velscale=10.0, probtol=0.2,
                       dyntol=0.05, dtmax=1e30, dtmin=0.0, jitter=None):
        """timestep.c:17-334; returns (new MaxPredTime array, number of clamped steps)"""
        class TS(C.Structure):
            _fields_ = [("crit", C.c_int), ("eta", C.c_double), ("velscale", C.c_double), ("probtol", C.c_double),
                        ("dyntol", C.c_double), ("dtmax", C.c_double), ("dtmin", C.c_double)]
        ts = TS(int(crit), eta, velscale, probtol, dyntol, dtmax, dtmin)
        active = np.ascontiguousarray(active, np.int32)
        accel = np.ascontiguousarray(accel, np.float32); curtime = np.ascontiguousarray(curtime, np.float32)
        mp = np.ascontiguousarray(maxpred, np.float32).copy()
        jit = np.zeros(len(active)) if jitter is None else np.ascontiguousarray(jitter, np.float64)
        self.L.ofind_timesteps.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_double] + [C.c_void_p] * 6
        self.L.ofind_timesteps.restype = C.c_int
        nc = self.L.ofind_timesteps(C.byref(self.par), C.byref(ts), len(active), _p(active), int(mode), float(time), float(vmax),
                                    _p(accel), _p(curtime), _p(mp), _p(self.hsml), _p(self.mass), _p(jit))
        return mp, nc

    def reflect(self, active, radius, pos, vel):
        """reflection.c:7-33; returns (new velocities, number reflected)"""
        active = np.ascontiguousarray(active, np.int32)
        pos = np.ascontiguousarray(pos, np.float32); v = np.ascontiguousarray(vel, np.float32).copy()
        self.L.oreflect.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        self.L.oreflect.restype = C.c_int
        n = self.L.oreflect(len(active), _p(active), float(radius), _p(pos), _p(v))
        return v, n

    def sidm_ensure_neighbours(self, dt, vmax):
        dta = np.ascontiguousarray(np.broadcast_to(np.float32(dt), (self.n,)), np.float32)
        return self.L.osidm_ensure(self.tree, C.byref(self.par), self.n, _p(self.vel), _p(self.mass), _p(self.hsml),
                                   _p(dta), _p(self.dvel), _p(self.ngb), _p(self.left), _p(self.right), float(vmax),
                                   self.rng)


class OracleForest:
    """Several collisionless particle types: one tree per type present (forcetree.c:90-158), every target walks all of
    them in ascending type order into the same accumulators with epsilon = max(eps of the tree's type, eps of the
    target's type) (forcetree.c:798-808 forces, :1397-1409 potentials).  Each tree is the single-type oracle tree of
    that type's particles in index order (the reference inserts them in that order)."""

    def __init__(self, pos, vel, mass, types, eps_table, **kw):
        self.pos = np.ascontiguousarray(pos, np.float32)
        self.types = np.ascontiguousarray(types, np.int32)
        self.eps = np.asarray(eps_table, np.float64)
        self.n = len(self.types)
        self.members, self.trees = {}, {}
        for t in sorted(set(self.types.tolist())):
            sel = np.nonzero(self.types == t)[0]
            self.members[t] = sel
            o = Oracle(self.pos[sel], np.ascontiguousarray(vel, np.float32)[sel], np.ascontiguousarray(mass, np.float32)[sel],
                       eps=float(self.eps[t]), **kw)
            o.treebuild()
            self.trees[t] = o
        L = lib()
        L.otree_force_at.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.otree_force_at.restype = None
        L.otree_potential_at.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.otree_potential_at.restype = None
        self.L = L

    def _walk(self, idx, oldacc, fn, out_width):
        idx = np.ascontiguousarray(idx, np.int32)
        oa_all = np.zeros(self.n, np.float32) if oldacc is None else np.ascontiguousarray(oldacc, np.float32)
        out = np.zeros((len(idx), out_width), np.float64)
        cost = np.zeros((len(idx), 2), np.int32)
        first = True
        for t, o in self.trees.items():                                   # ascending type order
            for s in sorted(set(self.types[idx].tolist())):               # targets grouped by their own type: one epsilon each
                rows = np.nonzero(self.types[idx] == s)[0]
                xyz = np.ascontiguousarray(self.pos[idx[rows]])
                oa = np.ascontiguousarray(oa_all[idx[rows]])
                a = np.ascontiguousarray(out[rows])
                c = np.ascontiguousarray(cost[rows])
                eps_save = o.par.eps
                o.par.eps = max(float(self.eps[t]), float(self.eps[s]))
                fn(o, len(rows), xyz, oa, a, c, 0 if first else 1)
                o.par.eps = eps_save
                out[rows] = a
                cost[rows] = c
            first = False
        return out, cost

    def force_tree(self, idx, oldacc=None):
        """double accelerations and (particle, node) interaction counts of force_treeevaluate() over all trees"""
        return self._walk(idx, oldacc, lambda o, n, xyz, oa, a, c, keep: self.L.otree_force_at(
            o.tree, C.byref(o.par), n, _p(xyz), _p(oa), _p(a), _p(c), keep), 3)

    def potential(self, idx, oldacc=None):
        """raw potentials as force_treeevaluate_potential() leaves them over all trees"""
        out, _ = self._walk(idx, oldacc, lambda o, n, xyz, oa, a, c, keep: self.L.otree_potential_at(
            o.tree, C.byref(o.par), n, _p(xyz), _p(oa), _p(a), keep), 1)
        return out[:, 0]

    def ngb_variable(self, i, h):
        """ngb_treefind_variable(P[i].PosPred, h, P[i].Type), forcetree.c:2163: neighbours of particle i inside its own
        type's tree, as global particle indices in the reference's list order"""
        t = int(self.types[i])
        lst, r2 = self.trees[t].ngb_variable(self.pos[i], h)
        return self.members[t][lst], r2
