"""The device-side schedule options (read from the environment when a context is created) must not
change results: mirrored y-tilde cache vs both copies updated (ffm.cpp:423-436, 451-464), the
3-kernel CG iteration vs separate direction / regulariser kernels, the fused same-side CG pass.
Checked in fp64 against the oracle and against each other."""
import importlib
import os

import numpy as np
import pytest

import ocffm
import pyoracle

pytestmark = pytest.mark.gpu

KNOBS = ["OCFFM_MIRROR_YT", "OCFFM_FUSED_DOT", "OCFFM_DIAG_FAST"]


def run(ds, prm, env):
    saved = {k: os.environ.get(k) for k in KNOBS}
    try:
        for k in KNOBS:
            os.environ.pop(k, None)
        os.environ.update(env)
        p = ocffm.Problem(ds, dtype=ocffm.F64, **prm)
    finally:
        for k, v in saved.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v
    model = p.init_model(seed=4)
    p.init_state()
    p.one_epoch()
    p.one_epoch()
    out = dict(cg=int(p.stats().cg_iters), obj=p.objective(), csr=p.vec("ytilde_csr"),
               csc=p.vec("ytilde_csc"), W={b: p.get_block(b[0], b[1], "W") for b in p.blocks()},
               val=p.validate())
    return model, out


@pytest.mark.parametrize("self_side", [True, False])
def test_schedule_knobs_do_not_change_results(self_side):
    synth = importlib.import_module("synth")
    ds = synth.generate("tiny", seed=9, test_rows=40, cold_rows=2)
    prm = dict(k=8, lam=0.5, omega=0.0625, r=-1.0, self_side=self_side, freq=True)
    model, base = run(ds, prm, {})
    o = pyoracle.Oracle(ds, **prm)
    for (f1, f2, which), w in model.items():
        o.set_block(f1, f2, which, w)
    o.init_state()
    o.one_epoch()
    o.one_epoch()
    assert base["cg"] == o.cg_iters_total()
    assert abs(base["obj"] - o.func()) <= 1e-9 * abs(o.func())
    # both orientations of the cache hold the same numbers entry for entry
    for env in ({"OCFFM_MIRROR_YT": "0"}, {"OCFFM_FUSED_DOT": "0"}, {"OCFFM_DIAG_FAST": "0"},
                {"OCFFM_MIRROR_YT": "0", "OCFFM_FUSED_DOT": "0", "OCFFM_DIAG_FAST": "0"}):
        _, alt = run(ds, prm, env)
        assert alt["cg"] == base["cg"], env
        assert abs(alt["obj"] - base["obj"]) <= 1e-11 * abs(base["obj"]), env
        assert np.allclose(alt["csr"], base["csr"], rtol=1e-9, atol=1e-12), env
        assert np.allclose(alt["csc"], base["csc"], rtol=1e-9, atol=1e-12), env
        for b, w in base["W"].items():
            assert np.allclose(alt["W"][b], w, rtol=1e-8, atol=1e-12), (env, b)
        assert np.array_equal(alt["val"]["topk"], base["val"]["topk"]), env
    # sums of the two orientations agree exactly up to ordering
    assert abs(float(np.sum(base["csr"])) - float(np.sum(base["csc"]))) <= 1e-9 * abs(float(np.sum(base["csr"])))
    assert np.array_equal(np.sort(base["csr"]), np.sort(base["csc"]))
