"""CPU: the oracle restatement against the unmodified reference run live (oracle/_ref), on a
larger seeded halo than the golden fixture.  Skipped where oracle/_ref is absent."""
import os
import tempfile

import numpy as np
import pytest

N = 12000
SIGMA = 41.78


@pytest.fixture(scope="module")
def pair(refdrv_mod):
    import oracle
    from sidm_b200 import ic
    pos, vel, mass, ids = ic.nfw(N, seed=21)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    R = refdrv_mod.Reference("diag")
    R.setup(N, CrossSectionInternal=SIGMA)
    R.init_rand(55)
    R.set_particles(pos, vel, mass, ids)
    R.treebuild()
    O = oracle.Oracle(pos, vel, mass, sigma=SIGMA)
    O.init_rand(55)
    O.treebuild()
    yield R, O, pos
    os.chdir(cwd)


def test_tree_and_forces(pair):
    R, O, pos = pair
    nd, od = R.dump_nodes(), O.dump()
    for k in ("center", "len", "mass", "s", "oc", "bmax2", "count"):
        assert np.array_equal(nd[k], od[k]), k
    assert np.array_equal(np.concatenate([nd["Q"], nd["P"][:, None]], axis=1), od["Q"])
    idx = np.arange(0, N, 23, dtype=np.int32)
    a_r, c_r = R.force_tree(idx)
    a_o, c_o = O.force_tree(idx)
    assert np.array_equal(a_r, a_o) and np.array_equal(c_r, c_o)
    assert np.array_equal(R.force_direct(idx), O.force_direct(idx))


def test_sidm_step_and_repair(pair, refdrv_mod):
    R, O, pos = pair
    R.setup_smoothinglengths_sidm(30)
    h = R.get("HSML")
    rng = np.random.default_rng(1)
    h = (h * rng.choice(np.array([1, 1, 1, 0.85, 1.25, 0.5], np.float32), N)).astype(np.float32)
    R.set("HSML", h)
    O.hsml[:] = h
    dt = 0.04
    R.all_active(0.0, dt / 2)
    vm = R.getvmax()
    assert vm == O.getvmax()
    dt32 = np.float32(2 * (R.time - 0.0))
    R.sidm()
    res = O.sidm(np.arange(N, dtype=np.int32), dt32, vm)
    assert np.array_equal(R.get("DVEL"), O.dvel) and np.array_equal(R.get("NGB"), O.ngb)
    R.sidm_ensure_neighbours(0)
    it = O.sidm_ensure_neighbours(dt32, vm)
    assert it > 0
    assert np.array_equal(R.get("HSML"), O.hsml)
    assert np.array_equal(R.get("NGB"), O.ngb)
    assert np.array_equal(R.get("DVEL"), O.dvel)
    log = refdrv_mod.read_scatlog("sct_000.0")
    assert len(log) >= res["sct"][2] > 0


def test_potential(pair):
    """force_treeevaluate_potential() (forcetree.c:1389-1755, both criteria) and compute_potential() (potential.c:18)"""
    R, O, pos = pair
    idx = np.arange(0, N, 19, dtype=np.int32)
    R.set("OLDACC", np.zeros(N, np.float32))
    assert np.array_equal(R.potential(idx), O.potential(idx, None)[0])
    R.all_active(0.0, 0.0)
    R.getvmax()
    R.compute_accelerations(1)
    old = R.get("OLDACC")
    assert np.array_equal(R.potential(idx), O.potential(idx, old)[0])
    R.compute_potential()
    assert np.array_equal(R.get("POT"), O.potential(np.arange(N, dtype=np.int32), old)[1])


def test_global_quantities(pair):
    """compute_global_quantities_of_system() (global.c:18-135) after compute_potential(): every one of the 102
    doubles of SysState bit for bit, with one and with three particle types"""
    import oracle
    R, O, pos = pair
    R.compute_potential()
    for types in (np.ones(N, np.int32), np.random.default_rng(8).choice(np.array([1, 2, 4], np.int32), N)):
        R.set("TYPE", types)
        ref = R.global_quantities()
        assert len(ref) == 102
        got = oracle.global_quantities(R.get("POSPRED"), R.get("VELPRED"), R.get("MASS"), R.get("POT"), types)
        assert np.array_equal(got, ref)
        assert ref[0] > 0 and ref[1] > 0 and ref[2] < 0          # Mass, EnergyKin, EnergyPot
    R.set("TYPE", np.ones(N, np.int32))


def test_snapshot_file_bytes(pair, refdrv_mod):
    """savepositions() (io.c:16-590): the oracle's snapshot writer against the file the reference writes - one type
    with and without a MassTable entry, three types of which one has tabulated masses"""
    import oracle
    R, O, pos = pair
    out = tempfile.mkdtemp()
    R.all_active(0.0, 0.004)
    R.getvmax()
    R.compute_accelerations(0)                       # PosPred / VelPred half a step ahead of Pos / Vel
    pp, vp, ids, m = R.get("POSPRED"), R.get("VELPRED"), R.get("ID"), R.get("MASS")
    assert not np.array_equal(pp, R.get("POS"))
    three = np.random.default_rng(8).choice(np.array([1, 2, 4], np.int32), N)
    cases = [(np.ones(N, np.int32), None), (np.ones(N, np.int32), [0, float(m[0]), 0, 0, 0, 0]), (three, [0, 0, 0.25, 0, 0, 0])]
    for k, (types, mt) in enumerate(cases):
        R.set("TYPE", types)
        path = R.savepositions(k, out, mass_table=mt, hubble_param=0.7)
        want = open(path, "rb").read()
        got = oracle.snapshot_bytes(pp, vp, ids, m, types, time=R.time, mass_table=mt, hubble_param=0.7, omega0=R.cfg["Omega0"])
        assert got == want, f"case {k}: {len(got)} vs {len(want)} bytes"
        back = oracle.read_snapshot(path)
        assert back["npart"].sum() == N and np.array_equal(np.sort(back["ids"]), np.sort(ids))
    R.set("TYPE", np.ones(N, np.int32))


def test_snapshot_periodic_wrap(refdrv_mod):
    """io.c:275-283 under -DPERIODIC: positions outside [0, BoxSize] are wrapped in the file (float += double);
    comoving header (redshift = 1/a - 1)"""
    import oracle
    from sidm_b200 import ic
    if not refdrv_mod.available("periodic"):
        pytest.skip("periodic reference build absent")
    BOX, A = 100.0, 0.25
    pos, vel, mass, ids = ic.periodic_box(12, seed=4, box=BOX, vel_sigma=60.0)
    n = len(mass)
    rng = np.random.default_rng(2)
    pos = (pos + rng.choice(np.array([0, 0, 0, -BOX, BOX, 2 * BOX], np.float32), (n, 3))).astype(np.float32)   # images
    cwd = os.getcwd()
    out = tempfile.mkdtemp()
    os.chdir(out)
    try:
        R = refdrv_mod.Reference("periodic")
        R.setup(n, BoxSize=BOX, SofteningHalo=0.5, ComovingIntegrationOn=1, Omega0=0.3, OmegaLambda=0.7, Hubble=0.1, Time=A)
        R.set_particles(pos, vel, mass, ids)
        path = R.savepositions(7, out, mass_table=[0, float(mass[0]), 0, 0, 0, 0], hubble_param=0.7)
        want = open(path, "rb").read()
        pp = R.get("POSPRED")
        assert (pp < 0).any() and (pp > BOX).any()
        got = oracle.snapshot_bytes(pp, R.get("VELPRED"), ids, mass, None, time=R.time, mass_table=[0, float(mass[0]), 0, 0, 0, 0],
                                    box=BOX, omega0=0.3, omega_lambda=0.7, hubble_param=0.7, comoving=True, periodic=True)
        assert got == want
        back = oracle.read_snapshot(path)
        assert back["mass"] is None and back["pos"].min() >= 0 and back["pos"].max() <= BOX
    finally:
        os.chdir(cwd)
