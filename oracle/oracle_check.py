"""smoke-test helper (TEST INFRASTRUCTURE): checks one small GPU result against the oracle."""
import numpy as np

import oracle


def smoke_check(hp, pos, vel, mass, ids, idx, acc_gpu):
    O = oracle.Oracle(pos, vel, mass)
    O.treebuild()
    oa = hp.get("OldAcc")
    acc, _ = O.force_tree(idx, oa)
    err = float(np.sqrt(((acc_gpu - acc) ** 2).sum() / (acc ** 2).sum()))
    print(f"smoke: GPU walk vs oracle walk rel rms {err:.3e}")
    assert err < 1e-4, err
    h2 = hp.ngb_treefind(idx[:64], 30)
    ref = np.array([O.ngb_treefind(pos[i], 30) for i in idx[:64]], np.float32)
    assert np.array_equal(h2, ref), "k-NN distances differ from the oracle"
    print("smoke: k-NN distances bit-exact vs oracle")
