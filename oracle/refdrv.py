"""ctypes driver for oracle/_ref/libsidmref*.so - the UNMODIFIED reference built single-rank
by oracle/Makefile (see oracle/ref_harness.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package.

The reference keeps all state in C globals and never frees it, so one process can hold
one problem size per loaded library; use a fresh process (or a size <= the first) to
change MaxPart.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

FIELDS = dict(POS=(0, 3, np.float32), VEL=(1, 3, np.float32), MASS=(2, 1, np.float32), ID=(3, 1, np.int32),
              TYPE=(4, 1, np.int32), CURTIME=(5, 1, np.float32), MAXPRED=(6, 1, np.float32),
              POSPRED=(7, 3, np.float32), VELPRED=(8, 3, np.float32), ACCEL=(9, 3, np.float32),
              POT=(10, 1, np.float32), GRAVCOST=(11, 1, np.float32), OLDACC=(12, 1, np.float32),
              FORCEFLAG=(13, 1, np.int32), LEFT=(14, 1, np.float32), RIGHT=(15, 1, np.float32),
              NGB=(16, 1, np.int32), HSML=(17, 1, np.float32), DVEL=(18, 3, np.float32))


class RefCfg(C.Structure):
    _fields_ = [("MaxPart", C.c_int), ("BufferSizeMB", C.c_int), ("TreeAllocFactor", C.c_double),
                ("ErrTolTheta", C.c_double), ("ErrTolForceAcc", C.c_double),
                ("TypeOfOpeningCriterion", C.c_int), ("ComovingIntegrationOn", C.c_int),
                ("MaxNodeMove", C.c_double), ("TreeUpdateFrequency", C.c_double), ("G", C.c_double),
                ("SofteningHalo", C.c_double), ("DesNumNgb", C.c_int), ("MaxNumNgbDeviation", C.c_int),
                ("CrossSectionInternal", C.c_double), ("ProbabilityTol", C.c_double),
                ("Seed1", C.c_int), ("Seed2", C.c_int), ("BoxSize", C.c_double),
                ("Omega0", C.c_double), ("OmegaLambda", C.c_double), ("Hubble", C.c_double),
                ("Time", C.c_double), ("YukawaVelocity", C.c_double), ("CrossSectionPowLaw", C.c_double),
                ("CrossSectionVelScale", C.c_double)]


DEFAULTS = dict(BufferSizeMB=100, TreeAllocFactor=0.8, ErrTolTheta=0.5, ErrTolForceAcc=0.005,
                TypeOfOpeningCriterion=1, ComovingIntegrationOn=0, MaxNodeMove=0.02,
                TreeUpdateFrequency=0.0, G=43007.1, SofteningHalo=0.3, DesNumNgb=30,
                MaxNumNgbDeviation=2, CrossSectionInternal=2.089, ProbabilityTol=0.2,
                Seed1=55, Seed2=497527, BoxSize=0.0, Omega0=1.0, OmegaLambda=0.0, Hubble=0.1,
                Time=0.0, YukawaVelocity=0.0, CrossSectionPowLaw=0.0, CrossSectionVelScale=1.0)


def lib_path(kind="diag"):
    name = {"diag": "libsidmref.so", "fast": "libsidmref_fast.so", "periodic": "libsidmref_per.so",
            "b200": "libsidmref_b200.so", "b200f": "libsidmref_b200f.so", "x1": "libsidmref_x1.so", "x2": "libsidmref_x2.so", "x4": "libsidmref_x4.so",
            "x3": "libsidmref_x3.so"}[kind]
    return os.path.join(HERE, "_ref", name)


def available(kind="diag"):
    return os.path.exists(lib_path(kind))


class Reference:
    """One loaded copy of the reference.  `kind`: 'diag' (parity build, -DDIAG -DSCATTERLOG
    -DFINDNBRLOG), 'fast' (shipped flags, for timing) or 'periodic' (-DPERIODIC)."""

    def __init__(self, kind="diag"):
        self.kind = kind
        self.lib = C.CDLL(lib_path(kind))
        L = self.lib
        L.ref_get_time.restype = C.c_double
        L.ref_getvmax.restype = C.c_double
        L.ref_get_vmax_global.restype = C.c_double
        if kind not in ("b200", "b200f"):   # tree / search accessors need the reference's own forcetree.c statics
            L.ref_ngb_treefind.restype = C.c_float
            L.ref_ngb_treefind.argtypes = [C.c_void_p, C.c_int, C.c_float]
            L.ref_ngb_variable.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_set_time.argtypes = [C.c_double]
        L.ref_set_vmax.argtypes = [C.c_double]
        L.ref_all_active.argtypes = [C.c_double, C.c_double]
        L.oracle_rng_log_begin.argtypes = [C.c_void_p, C.c_long]
        L.oracle_rng_log_count.restype = C.c_long
        L.oracle_rng_total_draws.restype = C.c_long
        self.n = 0
        self.psize = L.ref_sizeof_particle()

    # -- set-up -------------------------------------------------------------
    def setup(self, maxpart, **kw):
        cfg = dict(DEFAULTS)
        cfg.update(kw)
        c = RefCfg(MaxPart=int(maxpart), **cfg)
        rc = self.lib.ref_setup(C.byref(c))
        if rc != 0:
            raise RuntimeError("reference already set up with a smaller MaxPart")
        self.cfg = cfg

    def init_rand(self, seed):
        self.lib.ref_init_rand(int(seed))

    def set_particles(self, pos, vel, mass, ids):
        pos = np.ascontiguousarray(pos, np.float32)
        vel = np.ascontiguousarray(vel, np.float32)
        mass = np.ascontiguousarray(mass, np.float32)
        ids = np.ascontiguousarray(ids, np.int32)
        self.n = len(mass)
        self.lib.ref_set_particles(self.n, pos.ctypes, vel.ctypes, mass.ctypes, ids.ctypes)

    def get(self, name):
        fid, w, dt = FIELDS[name]
        out = np.empty((self.n, w) if w > 1 else (self.n,), dt)
        self.lib.ref_get_field(fid, out.ctypes)
        return out

    def set(self, name, arr):
        fid, w, dt = FIELDS[name]
        a = np.ascontiguousarray(arr, dt)
        assert a.size == self.n * w
        self.lib.ref_set_field(fid, a.ctypes)

    def get_raw(self):
        out = np.empty(self.n * self.psize, np.uint8)
        self.lib.ref_get_particles_raw(out.ctypes)
        return out

    def set_raw(self, raw):
        raw = np.ascontiguousarray(raw, np.uint8)
        self.n = raw.size // self.psize
        self.lib.ref_set_particles_raw(raw.ctypes, self.n)

    # -- state --------------------------------------------------------------
    @property
    def time(self):
        return self.lib.ref_get_time()

    def set_time(self, t):
        self.lib.ref_set_time(float(t))

    def getvmax(self):
        return self.lib.ref_getvmax()

    def set_vmax(self, v):
        self.lib.ref_set_vmax(float(v))

    def all_active(self, tcur, tnext):
        self.lib.ref_all_active(float(tcur), float(tnext))

    def active(self):
        n = self.lib.ref_num_active()
        out = np.empty(n, np.int32)
        self.lib.ref_get_active(out.ctypes)
        return out

    def set_active(self, idx):
        idx = np.ascontiguousarray(idx, np.int32)
        self.lib.ref_set_active(idx.ctypes, len(idx))

    def domain(self):
        mn = np.empty(3, np.float32)
        mx = np.empty(3, np.float32)
        self.lib.ref_get_domain(mn.ctypes, mx.ctypes)
        return mn, mx

    def cpu(self):
        out = np.zeros(7)
        self.lib.ref_get_cpu(out.ctypes)
        return dict(zip(["Gravity", "TreeConstruction", "TreeWalk", "EnsureNgb", "CommSum", "Predict", "TimeLine"], out))

    # -- tree ---------------------------------------------------------------
    def treebuild(self):
        return self.lib.ref_treebuild()

    def dump_nodes(self):
        nn = self.lib.ref_tree_numnodes()
        f = np.empty((nn, 24), np.float32)
        ii = np.empty((nn, 13), np.int32)
        self.lib.ref_dump_nodes(f.ctypes, ii.ctypes)
        return dict(len=f[:, 0], mass=f[:, 1], s=f[:, 2:5], center=f[:, 5:8], Q=f[:, 8:14], P=f[:, 14],
                    vs=f[:, 15:18], oc=f[:, 19], bmax2=f[:, 21], suns=ii[:, :8], sibling=ii[:, 8],
                    partind=ii[:, 9], count=ii[:, 10], cost=ii[:, 11], father=ii[:, 12])

    def dump_next(self):
        out = np.empty(self.n, np.int32)
        self.lib.ref_dump_next(out.ctypes)
        return out

    # -- forces -------------------------------------------------------------
    def force_tree(self, idx, want_cost=True):
        idx = np.ascontiguousarray(idx, np.int32)
        acc = np.empty((len(idx), 3), np.float64)
        cost = np.zeros((len(idx), 2), np.int32)
        self.lib.ref_force_tree(len(idx), idx.ctypes, acc.ctypes, cost.ctypes if want_cost else None)
        return acc, cost

    def potential(self, idx):
        idx = np.ascontiguousarray(idx, np.int32)
        pot = np.empty(len(idx), np.float64)
        self.lib.ref_potential(len(idx), idx.ctypes, pot.ctypes)
        return pot

    def compute_potential(self):
        self.lib.ref_compute_potential()

    def force_direct(self, idx):
        idx = np.ascontiguousarray(idx, np.int32)
        acc = np.empty((len(idx), 3), np.float64)
        self.lib.ref_force_direct(len(idx), idx.ctypes, acc.ctypes)
        return acc

    # -- neighbours ---------------------------------------------------------
    def ngb_variable(self, xyz, h, cap=4096):
        xyz = np.ascontiguousarray(xyz, np.float32)
        lst = np.empty(cap, np.int32)
        r2 = np.empty(cap, np.float32)
        n = self.lib.ref_ngb_variable(xyz.ctypes, C.c_float(h), lst.ctypes, r2.ctypes, cap)
        assert n <= cap
        return lst[:n].copy(), r2[:n].copy()

    def ngb_treefind(self, xyz, desngb, hguess=0.0):
        xyz = np.ascontiguousarray(xyz, np.float32)
        return float(self.lib.ref_ngb_treefind(xyz.ctypes, int(desngb), C.c_float(hguess)))

    # -- hot-path entry points ----------------------------------------------
    def gravity_tree(self):
        self.lib.ref_gravity_tree()

    def determine_interior(self):
        self.lib.ref_determine_interior()

    def sidm(self):
        self.lib.ref_sidm()

    def setup_nbr_sidm(self):
        self.lib.ref_setup_nbr_sidm()

    def sidm_ensure_neighbours(self, mode=0):
        self.lib.ref_sidm_ensure_neighbours(int(mode))

    def setup_smoothinglengths_sidm(self, desngb=30):
        self.lib.ref_setup_smoothinglengths_sidm(int(desngb))

    def compute_accelerations(self, mode=0):
        self.lib.ref_compute_accelerations(int(mode))

    def find_timesteps(self, mode, crit=0, eta=0.05, velscale=10.0, probtol=0.2, dyntol=0.05, dtmax=1e30, dtmin=0.0):
        self.lib.ref_find_timesteps.argtypes = [C.c_int, C.c_int] + [C.c_double] * 6
        self.lib.ref_find_timesteps(int(mode), int(crit), eta, velscale, probtol, dyntol, dtmax, dtmin)

    def reflect(self, radius):
        self.lib.ref_reflect.argtypes = [C.c_double]
        self.lib.ref_reflect(float(radius))

    def global_quantities(self):
        """compute_global_quantities_of_system(), global.c:18: SysState as a flat array of 102 doubles"""
        out = np.zeros(128, np.float64)
        n = self.lib.ref_global_quantities(out.ctypes)
        return out[:n]

    def savepositions(self, num, outdir, base="snap", mass_table=None, hubble_param=0.0):
        """savepositions(), io.c:16; returns the path of the file written"""
        mt = np.zeros(6) if mass_table is None else np.ascontiguousarray(mass_table, np.float64)
        d = os.path.join(outdir, "")
        self.lib.ref_savepositions.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_void_p, C.c_double]
        self.lib.ref_savepositions(int(num), d.encode(), base.encode(), mt.ctypes, float(hubble_param))
        return f"{d}{base}_{num:03d}"

    def run_steps(self, k):
        """k iterations of the main loop of run.c:34-150; returns (All.Time, NumForceUpdate) per iteration"""
        t = np.zeros(k, np.float64)
        na = np.zeros(k, np.int32)
        self.lib.ref_run_steps.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        self.lib.ref_run_steps(int(k), t.ctypes, na.ctypes)
        return t, na

    def set_softening(self, ptype, eps):
        self.lib.ref_set_softening.argtypes = [C.c_int, C.c_double]
        self.lib.ref_set_softening(int(ptype), float(eps))

    def advance(self):
        self.lib.ref_advance()
        return int(self.lib.ref_n_scat_particles())

    # -- RNG log ------------------------------------------------------------
    def rng_log_begin(self, cap):
        self._rngbuf = np.zeros(cap, np.float64)
        self.lib.oracle_rng_log_begin(self._rngbuf.ctypes, cap)

    def rng_log_end(self):
        n = self.lib.oracle_rng_log_count()
        self.lib.oracle_rng_log_end()
        assert n <= len(self._rngbuf), "rng log overflow"
        return self._rngbuf[:n].copy()


SCATLOG_DTYPE = np.dtype([("time", "<f4"), ("id1", "<i4"), ("id2", "<i4"), ("h1", "<f4"), ("h2", "<f4"),
                          ("x1", "<f4", 3), ("x2", "<f4", 3), ("v1", "<f4", 3), ("v2", "<f4", 3),
                          ("dv", "<f4", 3)])


def read_scatlog(path):
    """-DSCATTERLOG records written by the reference (sidm.h:1-10, sidm.c:571-601)."""
    if not os.path.exists(path):
        return np.zeros(0, SCATLOG_DTYPE)
    return np.fromfile(path, SCATLOG_DTYPE)
