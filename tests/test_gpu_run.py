"""GPU: a short integrated run through the C ABI - predict -> tree gravity (+ SIDM) -> advance, all on the
device - checked with the reference's own physics validation (energy_out / conservation):
total energy from compute_potential() is conserved by the leap-frog, momentum is conserved by the scatter kicks,
and elastic scattering conserves the kinetic energy of every pair."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _energy(hp, mass, t):
    n = len(mass)
    hp.predict_collisionless_only(t)                       # PosPred = Pos at a synchronised time
    pot = hp.compute_potential().astype(np.float64)
    vel = hp.peek("velh", np.float32, (n, 4))[:, :3].astype(np.float64)
    kin = 0.5 * (mass * (vel ** 2).sum(1)).sum()
    return kin, 0.5 * (mass * pot).sum()


def test_energy_conservation_gravity_only():
    from sidm_b200 import HotPath, ic
    n, dt, steps = 20000, 0.001, 60
    pos, vel, mass, ids = ic.hernquist(n, seed=41)
    m64 = mass.astype(np.float64)
    with HotPath(n, CrossSectionInternal=0.0) as hp:
        hp.set_particles(pos, vel, mass, ids)
        hp.predict_collisionless_only(0.0)
        hp.force_treebuild()
        hp.setup_smoothinglengths_sidm(30)
        hp.compute_accelerations(1, time=0.0, vmax=0.0)    # start-up forces: OldAcc for the relative criterion
        k0, w0 = _energy(hp, m64, 0.0)
        assert 0.35 < -k0 / w0 < 0.65, (k0, w0)            # the Eddington halo starts near virial equilibrium (2K = -W)
        t = 0.0
        for _ in range(steps):
            hp.compute_accelerations(0, time=t + dt / 2, vmax=0.0)
            hp.advance(time=t + dt / 2)
            t += dt
        k1, w1 = _energy(hp, m64, t)
    e0, e1 = k0 + w0, k1 + w1
    assert abs(e1 - e0) < 2e-3 * abs(e0), (e0, e1)


def test_scatter_conserves_momentum_and_pair_energy():
    from sidm_b200 import HotPath, ic
    n = 30000
    pos, vel, mass, ids = ic.hernquist(n, seed=42)
    with HotPath(n, CrossSectionInternal=150.0, Seed=3) as hp:
        hp.set_particles(pos, vel, mass, ids)
        hp.predict_collisionless_only(0.0)
        hp.force_treebuild()
        hp.setup_smoothinglengths_sidm(30)
        vmax = hp.getvmax()
        hp.set_particles(curtime=np.zeros(n, np.float32))
        hp.sidm(time=0.01, vmax=vmax)
        dv = hp.get("dVel").astype(np.float64)
        log = hp.scatlog()
        c = hp.counters()
    hit = np.abs(dv).sum(1) > 0
    assert c.sct_scattered == len(log) >= 50
    # pairs whose two members appear in no other event of this call keep exactly opposite kicks (equal masses:
    # momentum conserved); the reference's own "last writer wins" overwrites (sidm.c:527-529, 567-569) can only
    # touch particles that appear in more than one event
    i1, i2 = log["id1"] - 1, log["id2"] - 1
    ids_all, counts = np.unique(np.concatenate([i1, i2]), return_counts=True)
    multi = set(ids_all[counts > 1].tolist())
    clean = np.array([a not in multi and b not in multi for a, b in zip(i1, i2)])
    assert clean.sum() >= 40
    assert np.array_equal(dv[i1[clean]].astype(np.float32), log["dv"][clean])
    assert np.array_equal(dv[i2[clean]].astype(np.float32), -log["dv"][clean])
    assert hit.sum() == len(ids_all)
    # elastic and isotropic: |v1 - v2| is preserved by the logged kick (-DSCATTERLOG probe)
    v1, v2, d = log["v1"].astype(np.float64), log["v2"].astype(np.float64), log["dv"].astype(np.float64)
    before = np.sqrt(((v1 - v2) ** 2).sum(1))
    after = np.sqrt(((v1 + d - (v2 - d)) ** 2).sum(1))
    np.testing.assert_allclose(after, before, rtol=2e-6)
    h1 = log["h1"]
    r = np.sqrt(((log["x1"].astype(np.float64) - log["x2"]) ** 2).sum(1))
    assert (r < h1).all()                                  # partners lie inside the scatterer's kernel


def test_run_from_snapshot_with_device_statistics(tmp_path):
    """the reference's run loop with every per-step piece on the device: read_ic() of a format-1 file, start-up
    smoothing lengths and forces, 40 x [compute_accelerations(0) + advance()], energy statistics through
    compute_potential() + compute_global_quantities_of_system() (run.c:51-60 -> energy_out), savepositions().
    Checks: SysState energy conserved, linear momentum conserved by gravity + SIDM kicks, the statistics agree
    with a host evaluation of the same sums, the written snapshot reloads to the same state."""
    import oracle
    from sidm_b200 import HotPath, ic
    n, dt, steps = 20000, 0.001, 40
    pos, vel, mass, ids = ic.hernquist(n, seed=43)
    ic_file = str(tmp_path / "ic_000")
    open(ic_file, "wb").write(oracle.snapshot_bytes(pos, vel, ids, mass, None, time=0.0, mass_table=[0, float(mass[0]), 0, 0, 0, 0], omega0=1.0))
    with HotPath(n, CrossSectionInternal=20.89, Seed=5) as hp:
        t, mt, npart = hp.read_ic(ic_file)
        assert t == 0.0 and npart[1] == n
        hp.force_treebuild()
        hp.setup_smoothinglengths_sidm(30)
        vmax = hp.getvmax()
        hp.compute_accelerations(1, time=0.0, vmax=vmax)

        def stats(time):
            hp.predict_collisionless_only(time)
            pot = hp.compute_potential()
            return hp.compute_global_quantities_of_system(), pot
        s0, pot0 = stats(0.0)
        # the statistics against a host evaluation of the same sums
        pp, vp = hp.get("PosPred", "VelPred")
        want = oracle.global_quantities(pp, vp, mass, pot0)
        np.testing.assert_allclose(s0.flat()[:5], want[:5], rtol=1e-12)
        assert 0.35 < -s0.EnergyKin / s0.EnergyPot < 0.65
        scat = 0
        for _ in range(steps):
            hp.compute_accelerations(0, time=t + dt / 2, vmax=vmax)
            scat += hp.counters().sct_scattered
            hp.advance(time=t + dt / 2)
            t += dt
        s1, _ = stats(t)
        assert scat > 0, "fixture too quiet"
        assert abs(s1.EnergyTot - s0.EnergyTot) < 2e-3 * abs(s0.EnergyTot), (s0.EnergyTot, s1.EnergyTot)
        assert s1.Mass == s0.Mass
        pscale = float((mass.astype(np.float64) * np.sqrt((vel.astype(np.float64) ** 2).sum(1))).sum())
        for j in range(3):
            assert abs(s1.Momentum[j] - s0.Momentum[j]) < 1e-4 * pscale
        # snapshot of the end state, reloaded: the same predicted positions / velocities
        snap = str(tmp_path / "snap_001")
        hp.savepositions(snap, time=t, mass_table=mt)
        pp1, vp1 = hp.get("PosPred", "VelPred")
        hp.read_ic(snap)
        pp2, vp2 = hp.get("PosPred", "VelPred")
        assert np.array_equal(pp1, pp2) and np.array_equal(vp1, vp2)


def test_restart_is_bit_identical():
    """restart.c:37-154 writes All + P[] but not the generator state, so a restarted reference run scatters
    differently.  Here the generator state is two counters (b200_get_rng_state): a run stopped after 3 steps,
    dumped as particle arrays + those counters and continued in a fresh context equals the uninterrupted run
    bit for bit (positions, velocities, smoothing lengths, scatter kicks, the scatter log)."""
    from sidm_b200 import HotPath, ic
    n, dt = 30000, 0.002
    pos, vel, mass, ids = ic.hernquist(n, seed=44)
    kw = dict(CrossSectionInternal=150.0, Seed=11)

    def steps(hp, t, k, vmax):
        ev = []
        for _ in range(k):
            hp.compute_accelerations(0, time=t + dt / 2, vmax=vmax)
            log = hp.scatlog()
            ev.append(np.stack([log["id1"], log["id2"]], 1).astype(np.int64))
            hp.advance(time=t + dt / 2)
            t += dt
        return t, np.concatenate(ev)

    def dump(hp):
        velh = hp.peek("velh", np.float32, (n, 4))
        acc, oa, dv, ngb = hp.get("Accel", "OldAcc", "dVel", "NgbVelDisp")
        return dict(pos=hp.peek("pos0", np.float32, (n, 3)), vel=velh[:, :3].copy(), hsml=velh[:, 3].copy(), accel=acc, oldacc=oa,
                    dvel=dv, curtime=hp.peek("curtime", np.float32, (n,)), rng=hp.rng_state(), ngb=ngb)

    with HotPath(n, **kw) as hp:
        hp.set_particles(pos, vel, mass, ids)
        hp.predict_collisionless_only(0.0)
        hp.force_treebuild()
        hp.setup_smoothinglengths_sidm(30)
        vmax = hp.getvmax()
        hp.compute_accelerations(1, time=0.0, vmax=vmax)
        t, _ = steps(hp, 0.0, 3, vmax)
        rst = dump(hp)                                      # the "restart file"
        assert rst["rng"][0] >= 3
        t_end, ev_a = steps(hp, t, 3, vmax)
        end_a = dump(hp)
    assert len(ev_a) >= 20, "fixture too quiet"
    with HotPath(n, **kw) as hp:                            # fresh context, state from the dump
        hp.set_particles(rst["pos"], rst["vel"], mass, ids, curtime=rst["curtime"], accel=rst["accel"], oldacc=rst["oldacc"],
                         hsml=rst["hsml"], dvel=rst["dvel"])
        hp.set_field("ngb", rst["ngb"])
        hp.set_rng_state(rst["rng"])
        _, ev_b = steps(hp, t, 3, vmax)
        end_b = dump(hp)
        # control: without the generator state the continuation scatters other pairs
        hp.set_particles(rst["pos"], rst["vel"], mass, ids, curtime=rst["curtime"], accel=rst["accel"], oldacc=rst["oldacc"],
                         hsml=rst["hsml"], dvel=rst["dvel"])
        hp.set_field("ngb", rst["ngb"])
        hp.set_rng_state([0, 0])
        _, ev_c = steps(hp, t, 3, vmax)
    assert np.array_equal(ev_a, ev_b), "scattered pairs differ after the restart"
    for k in ("pos", "vel", "hsml", "accel", "oldacc", "dvel", "curtime", "ngb"):
        assert np.array_equal(end_a[k], end_b[k]), k
    assert np.array_equal(end_a["rng"], end_b["rng"])
    assert not (len(ev_c) == len(ev_a) and np.array_equal(ev_a, ev_c))
