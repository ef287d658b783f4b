"""Edge cases of the hot path on the GPU: latent sizes that need padding or a full warp per row,
empty / ragged rows, items never observed, and the C-ABI's error behaviour (never a crash)."""
import importlib

import numpy as np
import pytest

import ocffm
import pyoracle

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b)))) if a.size else 0.0


def pair(ds, prm, dtype, seed=3):
    o = pyoracle.Oracle(ds, **prm)
    p = ocffm.Problem(ds, dtype=dtype, **prm)
    for (f1, f2, which), w in p.init_model(seed=seed).items():
        o.set_block(f1, f2, which, w)
    o.init_state()
    p.init_state()
    return o, p


@pytest.mark.parametrize("k", [1, 5, 24, 64, 128])
@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_latent_sizes(k, dt):
    """k is padded to 4, 8, 32, 64, 128 on the device (lane groups of 1..32 lanes per row)."""
    synth = importlib.import_module("synth")
    dtype, tol = (ocffm.F64, 1e-9) if dt == "f64" else (ocffm.F32, 2e-4)
    ds = synth.generate("tiny", seed=5, test_rows=30, cold_rows=2)
    prm = dict(k=k, lam=0.5, omega=0.0625, r=-1.0, self_side=True, freq=False)
    o, p = pair(ds, prm, dtype)
    fu = p.fu
    for (f1, f2, which) in [(0, fu, "W"), (1, fu + 1, "H"), (0, 1, "H"), (fu, fu, "W")]:
        G = o.grad(f1, f2, which)
        assert rel_err(p.grad(f1, f2, which), G) <= tol
        assert rel_err(p.hess_vec(f1, f2, which, -G), o.hess_vec(f1, f2, which, -G)) <= tol
    o.one_epoch()
    p.one_epoch()
    if dt == "f64":
        assert int(p.stats().cg_iters) == o.cg_iters_total()
        assert abs(p.objective() - o.func()) <= 1e-9 * abs(o.func())
        for f1, f2 in p.blocks():
            assert rel_err(p.get_block(f1, f2, "W"), o.get_block(f1, f2, "W")) <= 1e-7
    ro, rp = o.validate(), p.validate()
    assert abs(rp["ploss"] - ro["ploss"]) <= 50 * tol * abs(ro["ploss"])
    if dt == "f64":
        assert np.array_equal(rp["topk"], ro["topk"])
        assert np.allclose(rp["ndcg"], ro["ndcg"], rtol=1e-9) and np.allclose(rp["prec"], ro["prec"], rtol=1e-12)


@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_empty_and_ragged_rows(dt):
    """Most users without any label, most items never observed (U->n < V->m), a few long rows."""
    synth = importlib.import_module("synth")
    dtype, tol = (ocffm.F64, 1e-9) if dt == "f64" else (ocffm.F32, 2e-4)
    ds = synth.generate("C1", seed=6, scale=0.05, test_rows=40, cold_rows=3, pos_override=0.4)
    lens = np.diff(ds.train.rowptr.astype(np.int64))
    assert (lens == 0).sum() > ds.m // 2 and ds.train.n_items <= ds.n
    prm = dict(k=8, lam=1.0, omega=0.03125, r=-1.0, self_side=True, freq=False)
    o, p = pair(ds, prm, dtype)
    for v in ("a", "b", "sa", "sb", "ytilde_csr", "ytilde_csc"):
        assert rel_err(p.vec(v), o.vec(v)) <= tol, v
    o.one_epoch()
    p.one_epoch()
    for v in ("a", "b", "ytilde_csr"):
        assert rel_err(p.vec(v), o.vec(v)) <= (1e-7 if dt == "f64" else 5e-2), v
    ro, rp = o.validate(), p.validate()
    if dt == "f64":
        assert np.array_equal(rp["topk"], ro["topk"])
        # fewer than 80 ranked items -> the tail of every list is padding, like the early `break`
        if ds.train.n_items < 80:
            assert np.all(rp["topk"][:, ds.train.n_items:] == 0xFFFFFFFF)


def test_error_behaviour_of_the_abi():
    synth = importlib.import_module("synth")
    ds = synth.generate("tiny", seed=5, test_rows=10)
    prm = dict(k=4, lam=0.5, omega=0.0625, r=-1.0)
    p = ocffm.Problem(ds, **prm)
    with pytest.raises(ocffm.OcffmError):      # solver before the model / state exist
        p.one_epoch()
    with pytest.raises(ocffm.OcffmError):
        p.init_state()                          # blocks never set
    p.init_model(seed=1)
    with pytest.raises(ocffm.OcffmError):
        p.grad(0, p.fu, "W")                    # init_state not called yet
    p.init_state()
    with pytest.raises(ocffm.OcffmError):
        p.grad(1, 0, "W")                       # f1 > f2
    with pytest.raises(AssertionError):
        p.set_block(0, p.fu, "W", np.zeros((3, 4)))   # wrong shape is caught before the ABI
    # a label beyond the item file: rejected (the reference would leave its two Y copies inconsistent)
    bad = synth.generate("tiny", seed=5)
    bad.train.idx = bad.train.idx.copy()
    bad.train.idx[0] = bad.n + 3
    with pytest.raises(ocffm.OcffmError):
        ocffm.Problem(bad, **prm)
    # --ns: same-side blocks do not exist
    q = ocffm.Problem(ds, self_side=False, **prm)
    with pytest.raises(ocffm.OcffmError):
        q.set_block(0, 0, "W", np.zeros((q.block_rows(0, 0, "W"), 4)))


def test_hyper_parameter_sweep_reuses_resident_data():
    """grid.sh-style sweep: one upload, several (lambda, omega) solves == fresh contexts (fp64)."""
    synth = importlib.import_module("synth")
    ds = synth.generate("C1", seed=9, scale=0.05, test_rows=40)
    base = dict(k=8, r=-1.0, self_side=True, freq=False)
    p = ocffm.Problem(ds, dtype=ocffm.F64, lam=1.0, omega=0.5, **base)
    model = p.init_model(seed=4)
    for lam, omega in [(4.0, 2.0 ** -7), (16.0, 2.0 ** -3), (1.0, 1.0)]:
        p.set_hyper(lam, omega, -1.0)
        for key, w in model.items():
            p.set_block(*key, w)
        p.init_state()
        p.one_epoch()
        fresh = ocffm.Problem(ds, dtype=ocffm.F64, lam=lam, omega=omega, **base)
        for key, w in model.items():
            fresh.set_block(*key, w)
        fresh.init_state()
        fresh.one_epoch()
        assert abs(p.objective() - fresh.objective()) <= 1e-11 * abs(fresh.objective())
        assert np.allclose(p.validate()["ndcg"], fresh.validate()["ndcg"], rtol=1e-9)
        fresh.close()


@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_device_side_model_init(dt):
    """ocffm_init_model: counter-based draw on the GPU -- reproducible from the seed,
    independent of the context, different per block / per W|H / per seed, uniform in
    (-0.1 qrsqrt(k), 0.1 qrsqrt(k)) like init_mat (ffm.cpp:3-12, 71-78); a solve from it matches the
    oracle fed with the same numbers."""
    synth = importlib.import_module("synth")
    dtype, tol = (ocffm.F64, 1e-9) if dt == "f64" else (ocffm.F32, 2e-4)
    ds = synth.generate("C1", seed=5, scale=0.1, test_rows=30)
    prm = dict(k=16, lam=4.0, omega=2.0 ** -7, r=-1.0, self_side=True, freq=False)
    p = ocffm.Problem(ds, dtype=dtype, **prm)
    p.init_model_device(seed=7)
    blocks = {(f1, f2, w): p.get_block(f1, f2, w) for f1, f2 in p.blocks() for w in "WH"}
    q = ocffm.Problem(ds, dtype=dtype, **prm)
    q.init_model_device(seed=7)
    for key, v in blocks.items():
        assert np.array_equal(q.get_block(*key), v), key          # same seed -> same model, any context
    q.init_model_device(seed=8)
    bound = 0.1 * 0.24957703567795358                            # 0.1 * qrsqrt(16) (probe value)
    allv = np.concatenate([v.ravel() for v in blocks.values()])
    assert np.all(np.abs(allv) <= bound * (1 + 1e-6)) and np.max(np.abs(allv)) > 0.98 * bound
    assert abs(float(np.mean(allv))) < 4 * bound / np.sqrt(3 * allv.size)         # mean of U(-b, b) within 4 sigma
    assert abs(float(np.var(allv)) - bound * bound / 3) < 0.02 * bound * bound / 3
    keys = sorted(blocks)
    for a_, b_ in zip(keys, keys[1:]):
        n = min(blocks[a_].size, blocks[b_].size)
        assert not np.array_equal(blocks[a_].ravel()[:n], blocks[b_].ravel()[:n])  # streams differ
    assert not np.array_equal(q.get_block(*keys[0]), blocks[keys[0]])              # seeds differ
    o = pyoracle.Oracle(ds, **prm)
    for key, v in blocks.items():
        o.set_block(*key, v)
    o.init_state()
    p.init_state()
    o.one_epoch()
    p.one_epoch()
    assert abs(p.objective() - o.func()) <= 10 * tol * abs(o.func())


def test_host_mirrors_are_current_after_every_outer_iteration():
    """ocffm_mirror_block: pinned fp64 mirrors are filled by one_epoch() itself (block by block on a
    second stream); they must equal what ocffm_get_block returns afterwards, bit for bit, for every
    iteration, and stay untouched once unregistered; pageable memory is refused."""
    import torch
    synth = importlib.import_module("synth")
    ds = synth.generate("C1", seed=5, scale=0.1, test_rows=30)
    p = ocffm.Problem(ds, dtype=ocffm.F32, k=16, lam=4.0, omega=2.0 ** -7, r=-1.0)
    p.init_model(seed=2)
    p.init_state()
    mirrors = {}
    for f1, f2 in p.blocks():
        for w in "WH":
            rows = p.block_rows(f1, f2, w)
            mirrors[(f1, f2, w)] = torch.zeros((rows, 16), dtype=torch.float64, pin_memory=True).numpy()
            p.mirror_block(f1, f2, w, mirrors[(f1, f2, w)])
    with pytest.raises(ocffm.OcffmError):
        p.mirror_block(0, 0, "W", np.zeros((p.block_rows(0, 0, "W"), 16)))      # not pinned
    for _ in range(2):
        p.one_epoch()
        for key, buf in mirrors.items():
            assert np.array_equal(buf, p.get_block(*key)), key
    key0 = sorted(mirrors)[0]
    p.mirror_block(*key0, None)
    before = mirrors[key0].copy()
    p.one_epoch()
    assert np.array_equal(mirrors[key0], before)                     # unregistered: left alone
    assert not np.array_equal(p.get_block(*key0), before)            # ... although the block moved on
    for key, buf in mirrors.items():
        if key != key0:
            assert np.array_equal(buf, p.get_block(*key)), key
