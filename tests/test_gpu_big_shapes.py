"""Parity on slices of the LARGE benchmark shapes (no parity run on C3/C4 shapes).

C4s: KDD12-shaped (fu=2, fv=4 -> 8 cross pairs, Kc = 256 at k=32, Zipf query field whose top
feature passes the DEFAULT hot-replica threshold).  C3s: Outbrain-shaped (no identity field at
all, Zipf user fields, k=16).  Gradient / Hessian-vector per phase from identical state, one
full outer iteration, and validate() (top-80 ids through the streaming-A tcgen05 scorer for
Kc = 256) against the oracle.  The oracle side is computed once per shape."""
import importlib

import numpy as np
import pytest

import ocffm
import pyoracle

pytestmark = pytest.mark.gpu
DT = {"f64": (ocffm.F64, 1e-9), "f32": (ocffm.F32, 1e-4)}
SHAPES = {"C4s": 32, "C3s": 16}


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    assert a.shape == b.shape
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b)))) if a.size else 0.0


def halves_of(fu, f):
    # a user-side and an item-side half of: first cross pair, last cross pair, a user side block, an item side block
    return [(0, fu, "W"), (0, fu, "H"), (fu - 1, f - 1, "W"), (fu - 1, f - 1, "H"), (0, 1, "W"), (1, 1, "H"),
            (fu, f - 1, "H"), (f - 1, f - 1, "W")]


@pytest.fixture(scope="module", params=sorted(SHAPES))
def case(request):
    synth = importlib.import_module("synth")
    shape, k = request.param, SHAPES[request.param]
    ds = synth.generate(shape, seed=3, test_rows=300, cold_rows=5)
    prm = dict(k=k, lam=4.0, omega=2.0 ** -7, r=-1.0, self_side=True, freq=False)
    o = pyoracle.Oracle(ds, **prm)
    rng = np.random.default_rng(11)
    blocks = {}
    for f1, f2 in o.blocks():
        for which in "WH":
            # the reference's init scale (init_mat, ffm.cpp:71-78: +-0.1/sqrt(k)).  Larger starts make the
            # first outer iteration chaotic on these Zipf-skewed shapes: with +-0.05 a 1e-15 relative
            # perturbation of the model moves the ORACLE's own objective after one iteration by 7e-5
            # (C3s) and its CG count by one, and the unmodified reference takes 190 vs 192 CG
            # iterations in its 2nd iteration with 1 vs 8 threads
            w = rng.uniform(-0.1 / np.sqrt(k), 0.1 / np.sqrt(k), size=(o.block_rows(f1, f2, which), k))
            blocks[(f1, f2, which)] = w
            o.set_block(f1, f2, which, w)
    o.init_state()
    fu, f = ds.users.f, ds.users.f + ds.items.f
    ref = dict(vec={v: o.vec(v) for v in ("a", "b", "sa", "sb", "ytilde_csr", "ytilde_csc")}, func0=o.func())
    ref["grad"], ref["hv"] = {}, {}
    for h in halves_of(fu, f):
        G = o.grad(*h)
        ref["grad"][h], ref["hv"][h] = G, o.hess_vec(*h, -G)
    o.one_epoch()
    ref["cg"] = o.cg_iters_total()
    ref["func1"] = o.func()
    ref["final"] = {(f1, f2, w): o.get_block(f1, f2, w) for f1, f2 in o.blocks() for w in "WH"}
    ref["vec1"] = {v: o.vec(v) for v in ("a", "b", "ytilde_csr", "ytilde_csc")}
    ref["val"] = o.validate(want_topk=True, want_scores=True)
    o.close()
    return shape, ds, prm, blocks, ref


@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_big_shape_slice_against_oracle(case, dt):
    shape, ds, prm, blocks, ref = case
    dtype, tol = DT[dt]
    p = ocffm.Problem(ds, dtype=dtype, **prm)        # default OCFFM_HOT_MIN: the Zipf field has hot features
    for key, w in blocks.items():
        p.set_block(*key, w)
    p.init_state()
    for v, want in ref["vec"].items():
        assert rel_err(p.vec(v), want) <= tol, v
    assert abs(p.objective() - ref["func0"]) <= max(tol, 1e-9) * abs(ref["func0"])
    for h, G in ref["grad"].items():
        assert rel_err(p.grad(*h), G) <= tol, ("grad", h)
        assert rel_err(p.hess_vec(*h, -G), ref["hv"][h]) <= tol, ("hv", h)
    p.reset_stats()
    p.one_epoch()
    cg = int(p.stats().cg_iters)
    # A whole outer iteration on these Zipf-skewed shapes is CHAOTIC in the summation order: the same
    # fp64 GPU code under different atomic / chunk orders (OCFFM_CHUNK, OCFFM_HOT_MIN, OCFFM_MIRROR_YT ...)
    # spreads by 5e-7 .. 1e-3 relative in the objective after the blocks of one iteration
    # (profiles/r02_variants_*.txt), CG stop tests flip, and the unmodified reference itself takes
    # 190 vs 192 CG iterations with 1 vs 8 threads.  So element-wise agreement of the
    # final model is not a meaningful test here; the per-phase checks above are the parity evidence,
    # and the iteration as a whole must land within the spread.
    # fp32 drifts further than the fp64 spread on these stiff solves (several halves stop at the 20-iteration
    # cap): measured 0.8-1.3% in the objective after the first outer iteration of the C4-shaped slice, with
    # 175-177 CG iterations against the oracle's 172 -- the per-phase bounds above hold at 1e-4 all the same
    # CG count within the spread of the summation-order variants: 5% in fp64; fp32 measured 3-5 of 172 on
    # C4s (stiff 20-iteration solves), bound 10%
    assert abs(cg - ref["cg"]) <= max(3, ref["cg"] // (20 if dt == "f64" else 10)), (cg, ref["cg"])
    assert abs(p.objective() - ref["func1"]) <= (3e-3 if dt == "f64" else 2.5e-2) * abs(ref["func1"]), (cg, ref["cg"])
    if dt == "f64":      # (fp32: the drifted trajectory moves single entries of a by a quarter of the largest one)
        for v in ("a", "b"):
            assert rel_err(p.vec(v), ref["vec1"][v]) <= 0.2, v
    # ranking parity from the ORACLE's model (Kc = 256 on C4s: the streaming-A tcgen05 variant in fp32)
    for key, want in ref["final"].items():
        p.set_block(*key, want)
    res, ro = p.validate(), ref["val"]
    Z, mism = ro["Z"], 0
    for i in range(ds.test_users.rows):
        for rnk in range(80):
            g, w = int(res["topk"][i, rnk]), int(ro["topk"][i, rnk])
            if g != w:
                mism += 1
                assert g != 0xFFFFFFFF
                assert abs(Z[i, g] - Z[i, w]) <= (1e-9 if dt == "f64" else 2e-5) * max(1.0, abs(Z[i, w])), (i, rnk, g, w)
    assert mism <= 300 * 80 * 0.01, mism
    if mism == 0:
        assert rel_err(res["prec"], ro["prec"]) <= 1e-12 and rel_err(res["ndcg"], ro["ndcg"]) <= 1e-9
    assert abs(res["ploss"] - ro["ploss"]) <= 10 * tol * abs(ro["ploss"])
    p.close()
