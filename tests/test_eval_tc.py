"""The tcgen05 3xTF32 scorer (eval_tc.cu) against the SIMT FP32 scorer (eval.cu) and the oracle:
identical top-80 ids except across near-ties, identical metrics when no id differs."""
import os
import subprocess
import sys

import numpy as np
import pytest

import ocffm
import pyoracle

pytestmark = pytest.mark.gpu


def run_validate(tc: bool, ds, prm, blocks):
    os.environ["OCFFM_EVAL_TC"] = "1" if tc else "0"
    try:
        p = ocffm.Problem(ds, dtype=ocffm.F32, **prm)
        for (f1, f2, which), w in blocks.items():
            p.set_block(f1, f2, which, w)
        return p.validate()
    finally:
        os.environ.pop("OCFFM_EVAL_TC", None)


# Kc = fu*fv*kp: 64 and 128 take the resident-A variant of k_score_topk_tc, 256 (C1 at k=64 -- the
# C5 configuration -- and the 2x4-field C4-shaped slice at k=32 -- the north-star configuration)
# the streaming-A variant (SmemTC<false>)
@pytest.mark.parametrize("shape,scale,k,rows", [("C1", 0.3, 16, 700), ("C1", 1.0, 32, 300), ("C1", 1.0, 64, 300),
                                                ("C4s", 1.0, 32, 300), ("C4s", 1.0, 64, 200)])
def test_tc_scorer_matches_simt_and_oracle(shape, scale, k, rows):
    import synth
    ds = synth.generate(shape, seed=4, scale=scale, test_rows=rows, cold_rows=7)
    prm = dict(k=k, lam=4.0, omega=2.0 ** -7, r=-1.0, self_side=True, freq=False)
    o = pyoracle.Oracle(ds, **prm)
    rng = np.random.default_rng(0)
    blocks = {}
    for f1, f2 in o.blocks():
        for which in "WH":
            w = rng.uniform(-0.3, 0.3, size=(o.block_rows(f1, f2, which), k))
            blocks[(f1, f2, which)] = w
            o.set_block(f1, f2, which, w)
    ro = o.validate(want_topk=True, want_scores=True)
    r_tc, r_simt = run_validate(True, ds, prm, blocks), run_validate(False, ds, prm, blocks)
    Z = ro["Z"]
    for name, res in (("tc", r_tc), ("simt", r_simt)):
        mism = 0
        for i in range(rows):
            for rnk in range(80):
                g, w = int(res["topk"][i, rnk]), int(ro["topk"][i, rnk])
                if g != w:
                    mism += 1
                    assert g != 0xFFFFFFFF, (name, i, rnk)
                    assert abs(Z[i, g] - Z[i, w]) <= 2e-5 * max(1.0, abs(Z[i, w])), (name, i, rnk, g, w)
        assert mism <= rows * 80 * 0.01, (name, mism)
        if mism == 0:
            assert np.allclose(res["ndcg"], ro["ndcg"], rtol=1e-9) and np.allclose(res["prec"], ro["prec"], rtol=1e-12)
        assert abs(res["ploss"] - ro["ploss"]) <= 1e-3 * abs(ro["ploss"])
