"""GPU: three particle types against the committed golden vectors of the unmodified reference
(tests/golden/global3k.npz + types3k.npz, written by tests/golden/make_golden.py global): forces of the three trees with
both opening criteria and epsilon = max(eps_tree, eps_target), interaction counts, raw potentials, start-up smoothing
lengths and neighbour counts inside the particle's own type.  Needs neither /root/reference nor oracle/_ref."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel_rms(a, b):
    return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))


def test_three_types_golden():
    from sidm_b200 import HotPath
    g = dict(np.load(os.path.join(GOLD, "global3k.npz")))
    t = dict(np.load(os.path.join(GOLD, "types3k.npz")))
    n, idx = len(g["mass"]), t["idx"]
    with HotPath(n, CrossSectionInternal=0.0, SofteningTable=[float(e) for e in g["eps"]], ReferenceNgbOrder=1) as hp:
        hp.set_particles(g["pospred"], g["velpred"], g["mass"], g["ids"], oldacc=np.zeros(n, np.float32))
        hp.set_field("ptype", g["types"])
        hp.predict_collisionless_only(0.0)
        hp.force_treebuild()
        acc, cost = hp.force_treeevaluate(idx)                       # OldAcc = 0: BH criterion
        assert rel_rms(acc, t["acc_bh"]) < 2e-6
        assert (cost.sum(1) == t["cost_bh"].sum(1)).mean() > 0.995
        hp.set_particles(oldacc=g["oldacc"])
        acc, cost = hp.force_treeevaluate(idx)                       # relative criterion
        assert rel_rms(acc, t["acc_rel"]) < 2e-6
        assert (cost.sum(1) == t["cost_rel"].sum(1)).mean() > 0.995
        np.testing.assert_allclose(hp.force_treeevaluate_potential(idx), t["pot_raw"], rtol=3e-6)
        hp.setup_smoothinglengths_sidm(30)
        h, ngb = hp.get("HsmlVelDisp", "NgbVelDisp")
        assert np.array_equal(ngb, t["ngb"])
        assert (h == t["hsml"]).mean() > 0.999
        np.testing.assert_allclose(h, t["hsml"], rtol=3e-7)


def test_three_types_against_oracle_forest():
    """a larger three-type halo against the oracle forest (oracle.OracleForest, pinned bit for bit on the golden
    vectors above by tests/test_oracle_golden.py)"""
    import oracle
    from sidm_b200 import HotPath, ic
    n = 60000
    pos, vel, mass, ids = ic.hernquist(n, seed=61)
    rng = np.random.default_rng(6)
    types = rng.choice(np.array([1, 2, 4], np.int32), n, p=[0.6, 0.3, 0.1]).astype(np.int32)
    mass = (mass * np.where(types == 4, 3.0, 1.0)).astype(np.float32)
    eps = [0.0, 0.3, 0.5, 0.0, 0.12, 0.0]
    F = oracle.OracleForest(pos, vel, mass, types, eps)
    idx = np.arange(0, n, 23, dtype=np.int32)
    with HotPath(n, CrossSectionInternal=0.0, SofteningTable=eps) as hp:
        hp.set_particles(pos, vel, mass, ids)
        hp.set_field("ptype", types)
        hp.predict_collisionless_only(0.0)
        hp.force_treebuild()
        acc, cost = hp.force_treeevaluate(idx)
        ref, cref = F.force_tree(idx, None)
        assert rel_rms(acc, ref) < 2e-6
        assert (cost.sum(1) == cref.sum(1)).mean() > 0.999
        hp.gravity_tree()                                               # sets OldAcc: relative criterion from here on
        old = hp.get("OldAcc")
        acc, cost = hp.force_treeevaluate(idx)
        ref, cref = F.force_tree(idx, old)
        assert rel_rms(acc, ref) < 2e-6
        assert (cost.sum(1) == cref.sum(1)).mean() > 0.999
        np.testing.assert_allclose(hp.force_treeevaluate_potential(idx), F.potential(idx, old), rtol=3e-6)
        hp.setup_smoothinglengths_sidm(30)
        h, ngb = hp.get("HsmlVelDisp", "NgbVelDisp")
        for i in idx[::9]:
            lst, _ = F.ngb_variable(int(i), h[i])
            assert len(lst) == ngb[i]
