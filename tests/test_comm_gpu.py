"""The in-library exchange (`rbp_comm_*`, include/rbp.h) on ONE GPU: a communicator of world size 1 runs the very same kernels — records
pushed into the owner's region buffer, the region-aware resolve / sort / fold, rows pushed to peers, the fused integer all-reduce — against
the plain single-process path.  (World 2 and 8 are checked bit for bit by tools/nlhe_world_check.py on multi-GPU boxes: profiles/r2m, r2n.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def comm(rbp):
    from robopoker_b200.comm import Comm
    c = Comm.create(0, 1, Comm.unique_id(), device=0)
    yield c
    c.close()


def test_nlhe_exchange_world_of_one_equals_plain_step(rbp, comm):
    from robopoker_b200.nlhe import Nlhe
    a = Nlhe(batch=256, seed=4, table_slots=1 << 18).step(5)
    b = Nlhe(batch=256, seed=4, table_slots=1 << 18).attach_comm(comm).step(5)
    assert a.profile().tobytes() == b.profile().tobytes()
    ca, cb = a.counters(), b.counters()
    assert {k: ca[k] for k in ("epochs", "nodes", "infos", "updates", "rows")} == {k: cb[k] for k in ("epochs", "nodes", "infos", "updates", "rows")}
    ms = b.step_timed(2, flush_l2=False)
    assert ms[0] > 0 and ms[5] > 0 and ms[6] > 0                 # the exchange phases ran and were timed
    a.step(2)
    assert a.profile().tobytes() == b.profile().tobytes()
    b.close(), a.close()


def test_kmeans_fused_allreduce_world_of_one(rbp, comm):
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from lloyd_data import turn_histograms
    pts = turn_histograms(4000, seed=2)
    outs = []
    for attach in (False, True):
        g = rbp.lloyd.Layer(pts, 16)
        if attach:
            g.attach_comm(comm)
        g.init_centroids(1)
        g.init_bounds()
        steps = [g.step() for _ in range(4)]
        outs.append((steps[-1].drift.copy(), steps[-1].sizes.copy(), steps[-1].reassignment, g.future()[0].copy()))
        g.close()
    assert np.array_equal(outs[0][0].view(np.uint32), outs[1][0].view(np.uint32)) and np.array_equal(outs[0][1], outs[1][1])
    assert outs[0][2] == outs[1][2] and np.array_equal(outs[0][3], outs[1][3])


def test_small_game_attach_needs_the_batched_fold(rbp, comm):
    g = rbp.Solver("leduc", batch=64, seed=1, fold=rbp.FOLD_BATCHED).attach_comm(comm).step(6)
    h = rbp.Solver("leduc", batch=64, seed=1, fold=rbp.FOLD_BATCHED).step(6)
    assert g.profile_rows().tobytes() == h.profile_rows().tobytes()
    with pytest.raises(rbp.RbpError):
        rbp.Solver("leduc", batch=64, seed=1).attach_comm(comm)   # the reference's ordered fold does not shard
