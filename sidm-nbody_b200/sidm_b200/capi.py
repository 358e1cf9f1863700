"""ctypes binding of include/sidm_b200.h (libsidm_b200.so) - the exact C ABI a maintainer of
the reference would bind; nothing here computes anything.

The product path is the CUDA library: if it is missing or there is no CUDA device every call
fails loudly (B200Error); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# B200_LIB: build-variant override for kernel experiments (scripts/); the default is the in-tree library
LIB_PATH = os.environ.get("B200_LIB") or os.path.join(PKG_DIR, "libsidm_b200.so")

ERRORS = {1: "tree nodes exhausted (forcetree.c:233 endrun(1))", 3: "allocation failed (endrun(3))",
          78: "neighbour list overflow (endrun(78))", 1155: "smoothing-length iteration failed (endrun(1155))",
          9001: "CUDA error", 9002: "no CUDA device - this library has no CPU path", 9003: "bad argument",
          9004: "bad state / not initialised", 9005: "coincident particles", 9007: "snapshot file could not be opened or written", 9006: "particle types changed behind the library's back"}


class B200Error(RuntimeError):
    def __init__(self, code, what=""):
        self.code = code
        super().__init__(f"libsidm_b200 {what}: error {code}: {ERRORS.get(code, '?')}")


class Params(C.Structure):
    _fields_ = [("device", C.c_int), ("MaxPart", C.c_int), ("TreeAllocFactor", C.c_double),
                ("ErrTolTheta", C.c_double), ("ErrTolForceAcc", C.c_double), ("TypeOfOpeningCriterion", C.c_int),
                ("ComovingIntegrationOn", C.c_int), ("G", C.c_double), ("SofteningTable", C.c_double * 6),
                ("BoxSize", C.c_double), ("PeriodicBoundariesOn", C.c_int),
                ("Omega0", C.c_double), ("OmegaLambda", C.c_double), ("Hubble", C.c_double),
                ("DesNumNgb", C.c_int), ("MaxNumNgbDeviation", C.c_int), ("CrossSectionInternal", C.c_double),
                ("CrossSectionType", C.c_int), ("YukawaVelocity", C.c_double), ("CrossSectionPowLaw", C.c_double),
                ("CrossSectionVelScale", C.c_double), ("Seed", C.c_ulonglong), ("BunchSizeSidm", C.c_int),
                ("ReferenceNgbOrder", C.c_int)]


class Layout(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("stride", "Pos", "Vel", "Mass", "ID", "Type", "CurrentTime", "PosPred",
                                       "VelPred", "Accel", "GravCost", "OldAcc", "Left", "Right", "NgbVelDisp",
                                       "HsmlVelDisp", "dVel", "MaxPredTime", "Potential")]


class TimestepParams(C.Structure):
    _fields_ = [("TypeOfTimestepCriterion", C.c_int), ("ErrTolIntAccuracy", C.c_double), ("ErrTolVelScale", C.c_double),
                ("ProbabilityTol", C.c_double), ("ErrTolDynamicalAccuracy", C.c_double), ("MaxSizeTimestep", C.c_double),
                ("MinSizeTimestep", C.c_double)]


class SysState(C.Structure):
    """struct state_of_system, allvars.h:517-537 (102 doubles)"""
    _fields_ = [("Mass", C.c_double), ("EnergyKin", C.c_double), ("EnergyPot", C.c_double), ("EnergyInt", C.c_double),
                ("EnergyTot", C.c_double), ("Momentum", C.c_double * 4), ("AngMomentum", C.c_double * 4),
                ("CenterOfMass", C.c_double * 4), ("MassComp", C.c_double * 5), ("EnergyKinComp", C.c_double * 5),
                ("EnergyPotComp", C.c_double * 5), ("EnergyIntComp", C.c_double * 5), ("EnergyTotComp", C.c_double * 5),
                ("MomentumComp", (C.c_double * 4) * 5), ("AngMomentumComp", (C.c_double * 4) * 5),
                ("CenterOfMassComp", (C.c_double * 4) * 5)]

    def flat(self):
        return np.frombuffer(bytes(self), np.float64).copy()


class Replay(C.Structure):
    _fields_ = [("rand", C.c_void_p), ("dir", C.c_void_p), ("extra", C.c_void_p), ("extra_off", C.c_void_p)]


class Counters(C.Structure):
    _fields_ = [("num_nodes", C.c_int), ("max_level", C.c_int), ("part_interactions", C.c_longlong),
                ("node_interactions", C.c_longlong), ("list_nodes", C.c_longlong), ("list_parts", C.c_longlong),
                ("num_targets", C.c_longlong), ("num_lists", C.c_longlong), ("sct_ntot", C.c_int), ("sct_pass1", C.c_int),
                ("sct_scattered", C.c_int), ("sct_rejected", C.c_int), ("ngb_candidates", C.c_longlong),
                ("ensure_iterations", C.c_int), ("ensure_repaired", C.c_int), ("ms_upload", C.c_float), ("ms_predict", C.c_float),
                ("ms_build", C.c_float), ("ms_walk", C.c_float), ("ms_sidm", C.c_float), ("ms_ensure", C.c_float),
                ("ms_download", C.c_float), ("kernel_launches", C.c_longlong)]


SCATLOG_DTYPE = np.dtype([("time", "<f4"), ("id1", "<i4"), ("id2", "<i4"), ("h1", "<f4"), ("h2", "<f4"),
                          ("x1", "<f4", 3), ("x2", "<f4", 3), ("v1", "<f4", 3), ("v2", "<f4", 3), ("dv", "<f4", 3)])

# struct particle_data with the reference's shipped flags (-DSIDM): allvars.h:422-460, 124 bytes
PARTICLE_DTYPE = np.dtype([("Pos", "<f4", 3), ("Vel", "<f4", 3), ("Mass", "<f4"), ("ID", "<i4"), ("Type", "<i4"),
                           ("CurrentTime", "<f4"), ("MaxPredTime", "<f4"), ("PosPred", "<f4", 3),
                           ("VelPred", "<f4", 3), ("Accel", "<f4", 3), ("Potential", "<f4"), ("GravCost", "<f4"),
                           ("OldAcc", "<f4"), ("ForceFlag", "<i4"), ("Left", "<f4"), ("Right", "<f4"),
                           ("NgbVelDisp", "<i4"), ("HsmlVelDisp", "<f4"), ("dVel", "<f4", 3)])


def layout_of(dtype=PARTICLE_DTYPE):
    f = dtype.fields
    return Layout(stride=dtype.itemsize, **{k: f[k][1] for k in ("Pos", "Vel", "Mass", "ID", "Type", "CurrentTime",
                                                                 "PosPred", "VelPred", "Accel", "GravCost", "OldAcc",
                                                                 "Left", "Right", "NgbVelDisp", "HsmlVelDisp", "dVel", "MaxPredTime", "Potential")})


EXPORTS = ["b200_init", "b200_set_params", "b200_finalize", "b200_last_cuda_error", "b200_set_stream", "b200_set_option", "b200_set_shard", "b200_current_stream", "b200_version",
           "b200_bind_particles", "b200_upload", "b200_download", "b200_download_to", "b200_upload_shard", "b200_download_shard", "b200_upload_active", "b200_download_active", "b200_bind_rows", "b200_upload_rows", "b200_shard_buffers", "b200_device_count", "b200_advance", "b200_find_timesteps", "b200_reflect", "b200_set_field", "b200_compute_potential", "b200_compute_global_quantities", "b200_savepositions", "b200_savepositions_part", "b200_get_rng_state", "b200_set_rng_state", "b200_load_snapshot", "b200_potential_raw", "b200_set_soa", "b200_get_soa", "b200_predict",
           "b200_tree_build", "b200_gravity", "b200_sidm", "b200_setup_nbr_sidm", "b200_sidm_ensure_neighbours",
           "b200_setup_smoothinglengths_sidm", "b200_compute_accelerations", "b200_getvmax", "b200_ngb_treefind",
           "b200_direct", "b200_walk_raw", "b200_get_tree", "b200_ngb_lists", "b200_sidm_debug", "b200_get_scatlog",
           "b200_get_counters", "b200_device_buffer"]

_lib = None


def load():
    """dlopen the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(9002, f"{LIB_PATH} not built - run __graft_entry__.build()")
        _lib = C.CDLL(LIB_PATH)
        _lib.b200_version.restype = C.c_char_p
        _lib.b200_gravity.argtypes = [C.c_void_p, C.c_int, C.c_double]
        _lib.b200_sidm.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p]
        _lib.b200_sidm_ensure_neighbours.argtypes = [C.c_int, C.c_double, C.c_double, C.c_void_p]
        _lib.b200_compute_accelerations.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_double]
        _lib.b200_predict.argtypes = [C.c_double]
        _lib.b200_set_stream.argtypes = [C.c_void_p]
        _lib.b200_set_option.argtypes = [C.c_char_p, C.c_int]
        _lib.b200_current_stream.restype = C.c_void_p
        _lib.b200_bind_particles.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        _lib.b200_set_soa.argtypes = [C.c_int] + [C.c_void_p] * 9
        _lib.b200_get_soa.argtypes = [C.c_void_p] * 10
        _lib.b200_get_tree.argtypes = [C.c_void_p] * 10
        _lib.b200_direct.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.b200_walk_raw.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib.b200_ngb_treefind.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.b200_ngb_lists.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib.b200_sidm_debug.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.b200_get_scatlog.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.b200_setup_nbr_sidm.argtypes = [C.c_void_p, C.c_int]
        _lib.b200_device_buffer.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p]
        _lib.b200_getvmax.argtypes = [C.c_void_p]
        _lib.b200_set_shard.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]
        _lib.b200_advance.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        _lib.b200_find_timesteps.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.b200_reflect.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        _lib.b200_set_field.argtypes = [C.c_char_p, C.c_void_p, C.c_longlong]
        _lib.b200_compute_potential.argtypes = [C.c_void_p]
        _lib.b200_compute_global_quantities.argtypes = [C.c_void_p]
        _lib.b200_savepositions.argtypes = [C.c_char_p, C.c_double, C.c_void_p, C.c_double, C.c_void_p]
        _lib.b200_savepositions_part.argtypes = [C.c_char_p, C.c_double, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib.b200_upload_active.argtypes = [C.c_void_p, C.c_int]
        _lib.b200_download_active.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.b200_load_snapshot.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.b200_get_rng_state.argtypes = [C.c_void_p]
        _lib.b200_set_rng_state.argtypes = [C.c_void_p]
        _lib.b200_potential_raw.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.b200_download_to.argtypes = [C.c_void_p]
        _lib.b200_upload_shard.argtypes = [C.c_int, C.c_int, C.c_int]
        _lib.b200_download_shard.argtypes = [C.c_void_p, C.c_int, C.c_int]
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise B200Error(rc, what)


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)
