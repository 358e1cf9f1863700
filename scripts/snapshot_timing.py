"""savepositions() / read_ic() on the device at BASELINE size (N = 1e7): wall time of b200_savepositions and
b200_load_snapshot to / from a tmpfs file, next to the path the reference takes through the boundary (download of the
124-byte AoS, then io.c's host loop - here numpy slicing of the same fields, which is faster than that loop)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sidm-nbody_b200"))
from sidm_b200 import HotPath, capi, ic  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
out = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
pos, vel, mass, ids = ic.nfw(n, seed=3)
with HotPath(n) as hp:
    hp.set_particles(pos, vel, mass, ids)
    hp.predict_collisionless_only(0.0)
    path = os.path.join(out, "b200_snap_000")
    for rep in range(3):
        t0 = time.perf_counter()
        hp.savepositions(path, time=0.0)
        t_save = time.perf_counter() - t0
    nbytes = os.path.getsize(path)
    aos = np.zeros(n, capi.PARTICLE_DTYPE)
    hp.bind_particles(aos, pin=True)
    hp.upload()
    t0 = time.perf_counter()
    hp.download()
    t_down = time.perf_counter() - t0
    t0 = time.perf_counter()
    with open(path + "_host", "wb") as f:
        for blk in (aos["PosPred"], aos["VelPred"], aos["ID"], aos["Mass"]):
            b = np.ascontiguousarray(blk)
            f.write(np.array([b.nbytes], np.int32).tobytes()); f.write(b.tobytes()); f.write(np.array([b.nbytes], np.int32).tobytes())
    t_host = time.perf_counter() - t0
    for rep in range(2):
        t0 = time.perf_counter()
        hp.read_ic(path)
        t_load = time.perf_counter() - t0
    print(f"N={n}: file {nbytes / 1e6:.0f} MB; b200_savepositions {t_save * 1e3:.0f} ms ({nbytes / t_save / 1e9:.1f} GB/s); "
          f"AoS download {t_down * 1e3:.0f} ms + host block writer {t_host * 1e3:.0f} ms; b200_load_snapshot {t_load * 1e3:.0f} ms")
    os.remove(path); os.remove(path + "_host")
