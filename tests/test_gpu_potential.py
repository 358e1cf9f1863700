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
