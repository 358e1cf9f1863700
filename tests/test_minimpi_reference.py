"""CPU: the UNMODIFIED reference with its real main() (main.c refuses one task) on 2 and 4 tasks of this box, through the
mini-MPI of oracle/stubs/minimpi (fork + shared-memory mailboxes).  Pins the mini-MPI on the reference's own parallel code:
read_ic.c distribution, domain.c ORB + particle exchange, the hypercube exchanges of gravtree.c / sidm.c, timeline.c
reductions and the parallel snapshot writer of io.c all have to work for the two runs to end at the same state.  The two
decompositions build different per-task trees (forces differ at the tree's own error), hence tolerances."""
import os

import numpy as np
import pytest

import mpi_case

EXE = os.path.join(mpi_case.ROOT, "oracle", "_ref", "sidm_ref_mpi")


@pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/sidm_ref_mpi not built (needs /root/reference)")
def test_reference_main_on_two_and_four_tasks(tmp_path):
    runs = {}
    for nt in (2, 4):
        w = str(tmp_path / f"np{nt}")
        mpi_case.write_case(w, 4000, TimeMax=0.005)
        r = mpi_case.run_case(w, "sidm_ref_mpi", nt, timeout=600)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        assert "Number of processors MUST" not in r.stdout
        runs[nt] = mpi_case.last_snapshot(w)
    a, b = runs[2], runs[4]
    assert a["time"] == b["time"] == 0.005
    assert np.array_equal(a["ids"], b["ids"]) and len(a["ids"]) == 4000
    np.testing.assert_allclose(a["pos"], b["pos"], rtol=0, atol=2e-3)
    np.testing.assert_allclose(a["vel"], b["vel"], rtol=0, atol=0.5)
    assert np.abs(a["vel"] - b["vel"]).mean() < 5e-3
