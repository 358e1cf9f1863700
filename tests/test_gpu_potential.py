"""GPU: compute_potential() (potential.c:18) / force_treeevaluate_potential() (forcetree.c:1389-1755) through the
C ABI against the oracle, whose restatement is pinned bit-exact on the unmodified reference
(tests/test_oracle_vs_reference.py::test_potential).  Tolerance: the force walk's (float interactions, same
interaction sets), 2e-6 relative."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_potential_matches_oracle():
    import oracle
    from sidm_b200 import HotPath, ic
    n = 30000
    pos, vel, mass, ids = ic.hernquist(n, seed=5)
    O = oracle.Oracle(pos, vel, mass)
    O.treebuild()
    idx = np.arange(0, n, 11, dtype=np.int32)
    with HotPath(n) as hp:
        hp.set_particles(pos, vel, mass, ids)
        hp.predict_collisionless_only(0.0)
        hp.force_treebuild()
        raw_bh = hp.force_treeevaluate_potential(idx)                  # OldAcc = 0: BH criterion
        ref_bh, _ = O.potential(idx, None)
        np.testing.assert_allclose(raw_bh, ref_bh, rtol=2e-6)
        hp.gravity_tree()                                               # sets OldAcc
        old = hp.get("OldAcc")
        raw = hp.force_treeevaluate_potential(idx)
        ref, _ = O.potential(idx, old)
        np.testing.assert_allclose(raw, ref, rtol=2e-6)
        pot = hp.compute_potential()
        _, want = O.potential(np.arange(n, dtype=np.int32), old)
        np.testing.assert_allclose(pot, want, rtol=3e-6)
        assert pot.max() < 0
        # direct summation check (softened Newtonian pairs) on a few targets: the tree error of the potential
        G, eps = 43007.1, 0.3
        for i in idx[:5]:
            d = np.sqrt(((pos.astype(np.float64) - pos[i]) ** 2).sum(1))
            far = d >= 2.8 * eps
            direct = -G * (mass[far] / d[far]).sum()
            near = (~far) & (d > 0)
            u = d[near] / (2.8 * eps)
            wp = np.where(u <= 0.5, 16 / 3 * u**2 - 48 / 5 * u**4 + 32 / 5 * u**5 - 14 / 5,
                          1 / (15 * u) + 32 / 3 * u**2 - 16 * u**3 + 48 / 5 * u**4 - 32 / 15 * u**5 - 16 / 5)
            direct += G * (mass[near] / (2.8 * eps) * wp).sum()
            assert abs(pot[i] - direct) < 2e-3 * abs(direct)


def test_global_quantities_golden_and_oracle():
    """compute_global_quantities_of_system() (global.c:18-135) on the device against the reference's own SysState
    (tests/golden/global3k.npz: three particle types) and against the oracle on a larger halo.  Same float
    products, double sums in tree order instead of particle order: 1e-12 of the sum of magnitudes."""
    import os
    import oracle
    from sidm_b200 import HotPath
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "global3k.npz")))
    n = len(g["mass"])
    x, v, m = g["pospred"].astype(np.float64), g["velpred"].astype(np.float64), g["mass"].astype(np.float64)
    vn, xn = np.sqrt((v * v).sum(1)), np.sqrt((x * x).sum(1))
    scale = dict(Momentum=(m * vn).sum(), AngMomentum=(m * xn * vn).sum(), CenterOfMass=(m * xn).sum() / m.sum())

    def check_state(st, ref, pot_rtol):
        want = oracle_struct(ref)
        for k in ("Mass", "EnergyKin"):
            assert abs(getattr(st, k) - want[k]) <= 1e-12 * abs(want[k]), k
        assert abs(st.EnergyPot - want["EnergyPot"]) <= pot_rtol * abs(want["EnergyPot"])
        assert abs(st.EnergyTot - want["EnergyTot"]) <= pot_rtol * abs(want["EnergyPot"])
        assert st.EnergyInt == 0
        for t in range(5):
            assert abs(st.MassComp[t] - want["MassComp"][t]) <= 1e-12 * want["Mass"]
            assert abs(st.EnergyKinComp[t] - want["EnergyKinComp"][t]) <= 1e-12 * want["EnergyKin"]
            assert abs(st.EnergyPotComp[t] - want["EnergyPotComp"][t]) <= pot_rtol * abs(want["EnergyPot"])
            for k in ("Momentum", "AngMomentum", "CenterOfMass"):
                for j in range(4):
                    assert abs(getattr(st, k + "Comp")[t][j] - want[k + "Comp"][t][j]) <= 1e-11 * scale[k], (k, t, j)
        for k in ("Momentum", "AngMomentum", "CenterOfMass"):
            for j in range(4):
                assert abs(getattr(st, k)[j] - want[k][j]) <= 1e-11 * scale[k], (k, j)

    def oracle_struct(flat):
        o, out = 0, {}
        for k, shape in (("Mass", ()), ("EnergyKin", ()), ("EnergyPot", ()), ("EnergyInt", ()), ("EnergyTot", ()),
                         ("Momentum", (4,)), ("AngMomentum", (4,)), ("CenterOfMass", (4,)), ("MassComp", (5,)),
                         ("EnergyKinComp", (5,)), ("EnergyPotComp", (5,)), ("EnergyIntComp", (5,)), ("EnergyTotComp", (5,)),
                         ("MomentumComp", (5, 4)), ("AngMomentumComp", (5, 4)), ("CenterOfMassComp", (5, 4))):
            cnt = int(np.prod(shape)) if shape else 1
            out[k] = flat[o:o + cnt].reshape(shape) if shape else float(flat[o])
            o += cnt
        assert o == 102
        return out

    with HotPath(n, SofteningTable=[float(e) for e in g["eps"]]) as hp:
        hp.set_particles(g["pospred"], g["velpred"], g["mass"], np.arange(1, n + 1, dtype=np.int32), oldacc=g["oldacc"])
        hp.set_field("ptype", g["types"])
        hp.predict_collisionless_only(0.0)
        # (1) the reduction alone: the reference's own P[].Potential goes in
        hp.set_field("potential", g["pot"])
        st = hp.compute_global_quantities_of_system()
        assert np.array_equal(st.flat().shape, (102,))
        check_state(st, g["sys"], 1e-12)
        # (2) with the device's own compute_potential() (three trees): the potential walk's tolerance
        pot = hp.compute_potential()
        np.testing.assert_allclose(pot, g["pot"], rtol=3e-6)
        check_state(hp.compute_global_quantities_of_system(), g["sys"], 3e-6)
    # single type, larger halo, oracle as the checker; PosPred != Pos (predicted half a step)
    from sidm_b200 import ic
    n = 200000
    pos, vel, mass, ids = ic.hernquist(n, seed=9)
    with HotPath(n) as hp:
        hp.set_particles(pos, vel, mass, ids)
        hp.compute_accelerations(1, time=0.0, vmax=0.0)
        hp.predict_collisionless_only(0.01)
        pot = hp.compute_potential()
        pp, vp = hp.get("PosPred", "VelPred")
        assert not np.array_equal(pp, pos)
        x, v, m = pp.astype(np.float64), vp.astype(np.float64), mass.astype(np.float64)
        vn, xn = np.sqrt((v * v).sum(1)), np.sqrt((x * x).sum(1))
        scale = dict(Momentum=(m * vn).sum(), AngMomentum=(m * xn * vn).sum(), CenterOfMass=(m * xn).sum() / m.sum())
        check_state(hp.compute_global_quantities_of_system(), oracle.global_quantities(pp, vp, mass, pot), 1e-12)
