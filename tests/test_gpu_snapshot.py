"""GPU: savepositions() (io.c:16-590) - the GADGET format-1 snapshot written from the device state
through b200_savepositions, byte for byte against (a) the reference's own file of the golden three-type fixture
(SHA-256 in tests/golden/global3k.npz) and (b) the oracle's writer, which tests/test_oracle_vs_reference.py pins on
the files the unmodified reference writes (incl. the -DPERIODIC wrap)."""
import hashlib
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_snapshot_golden_three_types():
    from sidm_b200 import HotPath
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "global3k.npz")))
    n = len(g["mass"])
    out = tempfile.mkdtemp()
    with HotPath(n, Omega0=float(g["snap_omega0"])) as hp:
        hp.set_particles(g["pospred"], g["velpred"], g["mass"], g["ids"])
        hp.set_field("ptype", g["types"])
        hp.predict_collisionless_only(0.0)
        path = os.path.join(out, "snap_003")
        npart = hp.savepositions(path, time=float(g["snap_time"]), mass_table=g["snap_mass_table"], hubble_param=0.7)
        assert npart.tolist() == [int((g["types"] == t).sum()) for t in range(5)] + [0]
        raw = open(path, "rb").read()
        assert len(raw) == int(g["snap_len"])
        assert raw[:264] == g["snap_head"].tobytes()
        assert hashlib.sha256(raw).hexdigest() == str(g["snap_sha256"])
        # error behaviour: unwritable path (io.c:98-102), gas particles (not on this path)
        from sidm_b200.capi import B200Error
        with pytest.raises(B200Error) as e:
            hp.savepositions(os.path.join(out, "no_such_dir", "snap"), time=0.0)
        assert e.value.code == 9007
        ty = g["types"].copy(); ty[5] = 0
        hp.set_field("ptype", ty)
        with pytest.raises(B200Error) as e:
            hp.savepositions(path, time=0.0)
        assert e.value.code == 9003


def test_snapshot_mid_run_against_oracle():
    """3.2e6 particles of one type half a step into a run (PosPred != Pos; the 38 MB blocks cross the 32 MB staging
    chunks), then the same particles as four types with a type-5 remainder that the file leaves out"""
    import oracle
    from sidm_b200 import HotPath, ic
    n = 3200000
    pos, vel, mass, ids = ic.hernquist(n, seed=13)
    out = tempfile.mkdtemp()
    with HotPath(n) as hp:
        hp.set_particles(pos, vel, mass, ids)
        hp.compute_accelerations(1, time=0.0, vmax=0.0)
        hp.predict_collisionless_only(0.003)
        pp, vp = hp.get("PosPred", "VelPred")
        assert not np.array_equal(pp, pos)
        path = os.path.join(out, "snap_000")
        hp.savepositions(path, time=0.003, hubble_param=0.7)
        assert open(path, "rb").read() == oracle.snapshot_bytes(pp, vp, ids, mass, None, time=0.003, hubble_param=0.7, omega0=1.0)
        back = oracle.read_snapshot(path)
        assert np.array_equal(back["ids"], ids) and np.array_equal(back["pos"], pp) and np.array_equal(back["mass"], mass)
        types = np.random.default_rng(1).choice(np.array([1, 2, 3, 4, 5], np.int32), n).astype(np.int32)
        hp.set_field("ptype", types)
        mt = [0, 0, float(mass[0]), 0, 0.5, 0]
        npart = hp.savepositions(path, time=0.003, mass_table=mt, hubble_param=0.7)
        assert npart.sum() == (types != 5).sum()
        raw = open(path, "rb").read()
        assert raw == oracle.snapshot_bytes(pp, vp, ids, mass, types, time=0.003, mass_table=mt, hubble_param=0.7, omega0=1.0)
        # and back: read_ic() of that file (38 MB blocks through the 32 MB staging buffers), written again -> the same bytes
        t, mt2, npart2 = hp.read_ic(path)
        assert t == 0.003 and np.array_equal(npart2, npart) and np.array_equal(mt2, mt) and hp.n == npart.sum()
        path2 = os.path.join(out, "snap_001")
        hp.savepositions(path2, time=t, mass_table=mt2, hubble_param=0.7)
        assert open(path2, "rb").read() == raw


def test_snapshot_periodic_wrap():
    import oracle
    from sidm_b200 import HotPath, ic
    BOX, A = 100.0, 0.25
    pos, vel, mass, ids = ic.periodic_box(24, seed=4, box=BOX, vel_sigma=60.0)
    n = len(mass)
    rng = np.random.default_rng(2)
    pos = (pos + rng.choice(np.array([0, 0, 0, -BOX, BOX, 2 * BOX], np.float32), (n, 3))).astype(np.float32)
    out = tempfile.mkdtemp()
    kw = dict(BoxSize=BOX, PeriodicBoundariesOn=1, SofteningHalo=0.5, ComovingIntegrationOn=1, Omega0=0.3, OmegaLambda=0.7, Hubble=0.1)
    with HotPath(n, **kw) as hp:
        hp.set_particles(pos, vel, mass, ids, curtime=np.full(n, A, np.float32))
        hp.predict_collisionless_only(A)
        pp, vp = hp.get("PosPred", "VelPred")
        path = os.path.join(out, "snap_007")
        hp.savepositions(path, time=A, mass_table=[0, float(mass[0]), 0, 0, 0, 0], hubble_param=0.7)
        want = oracle.snapshot_bytes(pp, vp, ids, mass, None, time=A, mass_table=[0, float(mass[0]), 0, 0, 0, 0], box=BOX,
                                     omega0=0.3, omega_lambda=0.7, hubble_param=0.7, comoving=True, periodic=True)
        assert open(path, "rb").read() == want
        back = oracle.read_snapshot(path)
        assert back["mass"] is None and back["pos"].min() >= 0 and back["pos"].max() <= BOX


def test_read_ic_golden_three_types():
    """read_ic() + init() start-up state (read_ic.c:32-481, init.c:76-100) from the reference's snapshot of the golden
    fixture (the oracle writer reproduces that file, SHA-256 checked): types from the block ranges, masses from the
    MassTable or the mass block, PosPred = Pos, VelPred = Vel, zeroed accelerations; then a full step on that state"""
    import oracle
    from sidm_b200 import HotPath
    from sidm_b200.capi import B200Error
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "global3k.npz")))
    n = len(g["mass"])
    raw = oracle.snapshot_bytes(g["pospred"], g["velpred"], g["ids"], g["mass"], g["types"], time=float(g["snap_time"]),
                                mass_table=g["snap_mass_table"], hubble_param=0.7, omega0=float(g["snap_omega0"]))
    assert hashlib.sha256(raw).hexdigest() == str(g["snap_sha256"])
    out = tempfile.mkdtemp()
    path = os.path.join(out, "snap_003")
    open(path, "wb").write(raw)
    order = np.concatenate([np.nonzero(g["types"] == t)[0] for t in range(5)])
    with HotPath(n + 100, Omega0=float(g["snap_omega0"]), SofteningTable=[float(e) for e in g["eps"]]) as hp:
        t, mt, npart = hp.read_ic(path)
        assert t == float(g["snap_time"]) and np.array_equal(mt, g["snap_mass_table"]) and hp.n == n
        posm = hp.peek("posm", np.float32, (n, 4))
        velh = hp.peek("velh", np.float32, (n, 4))
        assert np.array_equal(posm[:, :3], g["pospred"][order]) and np.array_equal(velh[:, :3], g["velpred"][order])
        assert np.array_equal(hp.peek("pos0", np.float32, (n, 3)), posm[:, :3])
        assert np.array_equal(hp.peek("pid", np.int32, (n,)), g["ids"][order])
        ty = hp.peek("ptype", np.int32, (n,))
        assert np.array_equal(ty, g["types"][order])
        want_m = np.where(ty == 2, np.float32(g["snap_mass_table"][2]), g["mass"][order]).astype(np.float32)
        assert np.array_equal(posm[:, 3], want_m)
        pp, vp, acc, oa, gc, dv = hp.get("PosPred", "VelPred", "Accel", "OldAcc", "GravCost", "dVel")
        assert np.array_equal(pp, posm[:, :3]) and np.array_equal(vp, velh[:, :3])
        assert not acc.any() and not oa.any() and not dv.any() and (gc == 1).all() and not velh[:, 3].any()
        # the loaded state runs: start-up smoothing lengths and one full step over the three trees
        hp.force_treebuild()
        hp.setup_smoothinglengths_sidm(30)
        hp.compute_accelerations(1, time=t, vmax=hp.getvmax())
        assert np.isfinite(hp.get("Accel")).all() and hp.get("NgbVelDisp").min() >= 28
        # written again: the same file
        path2 = os.path.join(out, "snap_004")
        hp.savepositions(path2, time=t, mass_table=mt, hubble_param=0.7)
        assert open(path2, "rb").read() == raw
        # error behaviour: truncated file, missing file
        open(path, "wb").write(raw[:len(raw) // 2])
        with pytest.raises(B200Error) as e:
            hp.read_ic(path)
        assert e.value.code == 9007
        with pytest.raises(B200Error) as e:
            hp.read_ic(os.path.join(out, "absent"))
        assert e.value.code == 9007


def test_snapshot_edge_cases():
    """three particles; every type with a MassTable entry (no mass block in the file, masses come from the table on
    load); a file that holds type-5 particles (other codes write them; io.c's writer leaves them out)"""
    import oracle
    from sidm_b200 import HotPath
    rng = np.random.default_rng(3)
    out = tempfile.mkdtemp()
    for n, types, mt in ((3, np.array([1, 1, 1], np.int32), None),
                         (500, rng.choice(np.array([1, 2, 3, 4], np.int32), 500), [0, 0.5, 0.25, 2.0, 1.0, 0])):
        pos = rng.standard_normal((n, 3)).astype(np.float32); vel = rng.standard_normal((n, 3)).astype(np.float32)
        mass = rng.random(n).astype(np.float32); ids = np.arange(1, n + 1, dtype=np.int32)
        want = oracle.snapshot_bytes(pos, vel, ids, mass, types, time=0.5, mass_table=mt, omega0=1.0)
        with HotPath(n) as hp:
            hp.set_particles(pos, vel, mass, ids)
            hp.set_field("ptype", types)
            hp.predict_collisionless_only(0.0)
            path = os.path.join(out, f"snap_{n}")
            hp.savepositions(path, time=0.5, mass_table=mt)
            assert open(path, "rb").read() == want
            t, mt2, npart = hp.read_ic(path)
            assert hp.n == n and t == 0.5
            ty = hp.peek("ptype", np.int32, (n,))
            assert np.array_equal(ty, np.sort(types))
            if mt is not None:                                   # masses from the table
                assert np.array_equal(hp.peek("posm", np.float32, (n, 4))[:, 3], np.asarray(mt, np.float32)[ty])
            hp.savepositions(path + "b", time=0.5, mass_table=mt)
            assert open(path + "b", "rb").read() == want
    # a file with type-5 particles: header and blocks by hand (the oracle writer follows io.c and drops them)
    n1, n5 = 40, 24
    n = n1 + n5
    pos = rng.standard_normal((n, 3)).astype(np.float32); vel = rng.standard_normal((n, 3)).astype(np.float32)
    mass = rng.random(n).astype(np.float32); ids = np.arange(1, n + 1, dtype=np.int32)
    base = oracle.snapshot_bytes(pos[:n1], vel[:n1], ids[:n1], mass[:n1], None, time=0.25, omega0=1.0)
    hdr = bytearray(base[4:260])
    hdr[20:24] = np.array([n5], np.int32).tobytes(); hdr[96 + 20:96 + 24] = np.array([n5], np.int32).tobytes()   # npart[5], npartTotal[5]
    rec = lambda b: np.array([len(b)], np.int32).tobytes() + b + np.array([len(b)], np.int32).tobytes()
    raw = rec(bytes(hdr)) + rec(pos.tobytes()) + rec(vel.tobytes()) + rec(ids.tobytes()) + rec(mass.tobytes())
    path = os.path.join(out, "snap_t5")
    open(path, "wb").write(raw)
    with HotPath(n) as hp:
        t, mt2, npart = hp.read_ic(path)
        assert npart.tolist() == [0, n1, 0, 0, 0, n5] and hp.n == n
        assert np.array_equal(hp.peek("ptype", np.int32, (n,)), np.r_[np.ones(n1, np.int32), np.full(n5, 5, np.int32)])
        assert np.array_equal(hp.peek("posm", np.float32, (n, 4)), np.c_[pos, mass])
        hp.savepositions(path + "b", time=t)                    # written back without the type-5 particles, like io.c
        assert open(path + "b", "rb").read() == base


def test_snapshot_split_over_files():
    """NumFilesPerSnapshot > 1 (io.c:78-103, 127-160): every file holds a contiguous range of rows (the particles of one group of
    tasks) in type order, npart = the file's counts, npartTotal = the system's, num_files = F; byte for byte the oracle writer's
    file of the same rows; the files together hold every particle once"""
    import oracle
    from sidm_b200 import HotPath, ic
    n = 50000
    pos, vel, mass, ids = ic.hernquist(n, seed=23)
    types = np.random.default_rng(3).choice(np.array([1, 2, 4], np.int32), n, p=[0.6, 0.3, 0.1]).astype(np.int32)
    mt = [0, 0, float(mass[0]), 0, 0, 0]
    out = tempfile.mkdtemp()
    cuts = [0, 17001, 17001, 41234, n]                       # four files, one of them empty
    tot = [int((types == t).sum()) for t in range(5)] + [0]
    with HotPath(n, SofteningTable=[0, 0.3, 0.3, 0, 0.3, 0]) as hp:
        hp.set_particles(pos, vel, mass, ids)
        hp.set_field("ptype", types)
        hp.predict_collisionless_only(0.0)
        seen = []
        for f in range(4):
            a, b = cuts[f], cuts[f + 1]
            path = os.path.join(out, f"snap_007.{f}")
            npart = hp.savepositions_part(path, a, b - a, 4, time=0.25, mass_table=mt, hubble_param=0.7)
            assert npart.tolist() == [int((types[a:b] == t).sum()) for t in range(5)] + [0]
            ref = oracle.snapshot_bytes(pos[a:b], vel[a:b], ids[a:b], mass[a:b], types[a:b], time=0.25, mass_table=mt, hubble_param=0.7,
                                        omega0=1.0, npart_total=tot, num_files=4)
            assert open(path, "rb").read() == ref
            if b > a:
                seen.append(oracle.read_snapshot(path)["ids"])
        assert np.array_equal(np.sort(np.concatenate(seen)), np.sort(ids))


def test_read_ic_split_over_files():
    """read_ic() of initial conditions split over files (read_ic.c:62-75): path.0 .. path.3 written by the part writer, read back
    as one particle order; types, masses (mass block and MassTable), ids and positions as written"""
    from sidm_b200 import HotPath, ic
    n = 40000
    pos, vel, mass, ids = ic.hernquist(n, seed=29)
    types = np.random.default_rng(5).choice(np.array([1, 2, 3], np.int32), n, p=[0.5, 0.3, 0.2]).astype(np.int32)
    mass = mass.copy(); mass[types == 2] = np.float32(mass[0] * 2)
    mt = [0, 0, float(mass[0] * 2), 0, 0, 0]
    out = tempfile.mkdtemp()
    cuts = [0, 9000, 21000, 21000, n]
    base = os.path.join(out, "ics")
    with HotPath(n, SofteningTable=[0, 0.3, 0.3, 0.3, 0, 0]) as hp:
        hp.set_particles(pos, vel, mass, ids)
        hp.set_field("ptype", types)
        hp.predict_collisionless_only(0.0)
        for f in range(4):
            hp.savepositions_part(f"{base}.{f}", cuts[f], cuts[f + 1] - cuts[f], 4, time=0.5, mass_table=mt)
    with HotPath(n, SofteningTable=[0, 0.3, 0.3, 0.3, 0, 0]) as hp:
        t, mtab, npart = hp.read_ic(base)
        assert t == 0.5 and hp.n == n and npart.tolist() == [int((types == k).sum()) for k in range(5)] + [0]
        # expected order: file after file, inside a file type after type
        order = np.concatenate([np.arange(cuts[f], cuts[f + 1])[np.argsort(types[cuts[f]:cuts[f + 1]], kind="stable")] for f in range(4)])
        posm = hp.peek("posm", np.float32, (n, 4))
        assert np.array_equal(posm[:, :3], pos[order]) and np.array_equal(posm[:, 3], mass[order])
        assert np.array_equal(hp.peek("pid", np.int32, (n,)), ids[order]) and np.array_equal(hp.peek("ptype", np.int32, (n,)), types[order])
        assert np.array_equal(hp.peek("velh", np.float32, (n, 4))[:, :3], vel[order])
        hp.predict_collisionless_only(0.5)
        hp.force_treebuild()                                 # the loaded state is usable
