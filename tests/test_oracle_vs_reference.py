"""Pins the plain-C oracle (oracle/ocffm_oracle.c) to the UNMODIFIED reference.

Every array compared here was written by oracle/ref_harness.cpp calling the reference's own
functions (tests/golden/*.npz, minted by oracle/make_goldens.py) or by the reference's nDCG
known-answer tool (tests/golden/ndcg_case1.json).  Tolerance is fp64 summation-order noise.
"""
import json
import os

import numpy as np
import pytest

import pyoracle
from conftest import GOLDEN, params_of

RTOL = 1e-9


def make_oracle(d):
    prm, nr_pass = params_of(d)
    ds = pyoracle.dataset_from_dump(d)
    o = pyoracle.Oracle(ds, **prm)
    return o, ds, prm, nr_pass


def load_blocks(o, d, pfx):
    for f1, f2 in o.blocks():
        o.set_block(f1, f2, "W", d[f"{pfx}.{f1}_{f2}.W"])
        o.set_block(f1, f2, "H", d[f"{pfx}.{f1}_{f2}.H"])


def close(a, b, rtol=RTOL):
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    assert a.shape == b.shape
    scale = max(1e-300, float(np.max(np.abs(b))) if b.size else 1.0)
    assert float(np.max(np.abs(a - b))) <= rtol * scale if a.size else True


def test_csc_and_popular_match_reference_reader(golden):
    name, d = golden
    o, ds, _, _ = make_oracle(d)
    colptr, rowidx = o.csc()
    assert np.array_equal(colptr, d["V.Y.rowptr"])           # transY, ffm.cpp:259-294: bit-exact
    assert np.array_equal(rowidx, d["V.Y.idx"].astype(np.uint32))
    close(o.vec("popular"), d["U.popular"], 1e-15)
    assert ds.train.n_items == int(d["U.hdr"][1])


def test_init_state(golden):
    name, d = golden
    o, ds, prm, _ = make_oracle(d)
    load_blocks(o, d, "init")
    o.init_state()
    for f1, f2 in o.blocks():
        close(o.embed(f1, f2, "P"), d[f"init.{f1}_{f2}.P"])
        close(o.embed(f1, f2, "Q"), d[f"init.{f1}_{f2}.Q"])
    for v in ("a", "b", "sa", "sb", "ytilde_csr", "ytilde_csc"):
        close(o.vec(v), d["init." + v])
    if "init.func" in d:
        close(o.func(), d["init.func"])


def test_rng_init_matches_libstdcxx(golden):
    """init_mat (ffm.cpp:71-78): rand()-seeded minstd_rand0 + uniform_real_distribution."""
    name, d = golden
    import ctypes
    ctypes.CDLL(None).srand(1)   # the reference never seeds: glibc's default seed is 1
    o, ds, prm, _ = make_oracle(d)
    o.init_model_rng()
    for f1, f2 in o.blocks():
        for which in "WH":
            got, want = o.get_block(f1, f2, which).ravel(), d[f"init.{f1}_{f2}.{which}"]
            # allow 1-2 ulp: the reference binary may contract u*(b-a)+a into an FMA
            assert np.max(np.abs(got - want)) <= 4 * np.finfo(np.float64).eps * np.max(np.abs(want))


PROBES = ["side_u.W", "side_u.H", "side_v.W", "side_v.H", "cross.W", "cross.H", "cross_last.W", "cross_last.H"]


def probe_block(o, probe):
    fu, f = o.fu, o.f
    kind, which = probe.split(".")
    blk = {"side_u": (0, 0), "side_v": (fu, f - 1), "cross": (0, fu), "cross_last": (fu - 1, f - 1)}[kind]
    return blk[0], blk[1], which


@pytest.mark.parametrize("probe", PROBES)
def test_gradient_hessvec_cg(golden, probe):
    name, d = golden
    if f"probe.{probe}.G" not in d:
        pytest.skip("no same-side blocks under --ns")
    o, ds, prm, _ = make_oracle(d)
    load_blocks(o, d, "init")
    o.init_state()
    f1, f2, which = probe_block(o, probe)
    G = o.grad(f1, f2, which)
    close(G, d[f"probe.{probe}.G"])
    Hv = o.hess_vec(f1, f2, which, -d[f"probe.{probe}.G"].reshape(G.shape))
    close(Hv, d[f"probe.{probe}.Hv"])
    S, it = o.cg(f1, f2, which, d[f"probe.{probe}.G"].reshape(G.shape))
    assert it == int(d[f"probe.{probe}.cg_iters"][0])
    close(S, d[f"probe.{probe}.S"], 1e-8)


def test_epochs_and_validate(golden):
    name, d = golden
    o, ds, prm, nr_pass = make_oracle(d)
    load_blocks(o, d, "init")
    o.init_state()
    funcs, cgs, prev = [], [], 0
    for e in range(nr_pass):
        o.one_epoch()
        cgs.append(o.cg_iters_total() - prev)
        prev = o.cg_iters_total()
        if prm["self_side"]:
            funcs.append(o.func())
        if e == 0 and "epoch1.a" in d:
            for v in ("a", "b", "sa", "sb", "ytilde_csr", "ytilde_csc"):
                close(o.vec(v), d["epoch1." + v], 1e-8)
    assert cgs == [int(x) for x in d["epochs.cg_iters"]]
    if prm["self_side"]:
        close(funcs, d["epochs.func"], 1e-9)
    for f1, f2 in o.blocks():
        close(o.get_block(f1, f2, "W"), d[f"final.{f1}_{f2}.W"], 1e-7)
        close(o.get_block(f1, f2, "H"), d[f"final.{f1}_{f2}.H"], 1e-7)
    for v in ("a", "b", "ytilde_csr", "ytilde_csc"):
        close(o.vec(v), d["final." + v], 1e-7)

    res = o.validate(want_topk=True, want_scores=True)
    close(res["prec"], d["va.prec"], 1e-12)
    close(res["ndcg"], d["va.ndcg"], 1e-9)
    close(res["ploss"], d["va.ploss"], 1e-8)
    # raw scores of warm rows, and the top-80 they imply (first maximum wins ties)
    Zref = d["va.Z"]
    nnx = d["T.nnx"]
    un = int(d["U.hdr"][1])
    for i in range(Zref.shape[0]):
        if nnx[i] == 0:
            z = d["U.popular"].copy()
        else:
            close(res["Z"][i], Zref[i], 1e-8)
            z = Zref[i, :un].copy()
        order = []
        for _ in range(min(80, un)):
            j = int(np.argmax(z))
            order.append(j)
            z[j] = -1000.0
        got = res["topk"][i][:len(order)]
        # ids must agree wherever the reference's scores are not within rounding of a tie
        zs = (d["U.popular"] if nnx[i] == 0 else Zref[i, :un])
        for rnk, (g, w) in enumerate(zip(got, order)):
            assert g == w or abs(zs[g] - zs[w]) <= 1e-9 * max(1.0, abs(zs[w])), (i, rnk, g, w)


def test_ndcg_known_answer_fixture():
    """The only result the reference's own tooling pins (script/nDCG_degub_tool)."""
    with open(os.path.join(GOLDEN, "ndcg_case1.json")) as fh:
        fx = json.load(fh)
    for labels, want in zip(fx["labels"], fx["ndcg_at_10"]):
        got = pyoracle.ndcg_at(fx["ranking"], labels, 10)
        assert abs(got - want) < 1e-4, (labels, got, want)
