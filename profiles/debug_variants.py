"""fp64 per-block objective trajectories of one outer iteration under different summation orders
(environment knobs), to separate inherent CG sensitivity from bugs.  GPU only; the oracle's own
trajectory (computed on the builder's CPU) is read from profiles/debug_oracle_<shape>.json if present."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "one-class-ffm_b200"))
import numpy as np
import ocffm, synth

shape = sys.argv[1] if len(sys.argv) > 1 else "C4s"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 32
ds = synth.generate(shape, seed=3, test_rows=300, cold_rows=5)
prm = dict(k=k, lam=4.0, omega=2.0 ** -7, r=-1.0, self_side=True, freq=False)
fu = ds.users.f
VARIANTS = [{}, {"OCFFM_HOT_MIN": "0"}, {"OCFFM_CHUNK": "32"}, {"OCFFM_MIRROR_YT": "0"}, {"OCFFM_FUSED_DOT": "0"},
            {"OCFFM_DIAG_FAST": "0"}]
ref = None
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"debug_oracle_{shape}.json")
if os.path.exists(path):
    ref = json.load(open(path))
    print("oracle   ", " ".join(f"{v:.10e}" for v in ref["obj"]), ref["cg"])
for env in VARIANTS:
    for kk in ("OCFFM_HOT_MIN", "OCFFM_CHUNK", "OCFFM_MIRROR_YT", "OCFFM_FUSED_DOT", "OCFFM_DIAG_FAST"):
        os.environ.pop(kk, None)
    os.environ.update(env)
    p = ocffm.Problem(ds, dtype=ocffm.F64, **prm)
    rng = np.random.default_rng(11)
    sc = 0.1 / np.sqrt(k)
    blocks = sorted({(f1, f2) for f1, f2 in p.blocks()})
    for f1, f2 in blocks:
        for which in "WH":
            p.set_block(f1, f2, which, rng.uniform(-sc, sc, size=(p.block_rows(f1, f2, which), k)))
    p.init_state()
    order = sorted(blocks, key=lambda b: (2 if b[0] < fu <= b[1] else (0 if b[1] < fu else 1), b))
    objs, cgs = [], []
    for f1, f2 in order:
        p.reset_stats()
        p.solve_block(f1, f2)
        cgs.append(int(p.stats().cg_iters))
        objs.append(p.objective())
    print(f"{str(env):28s}", " ".join(f"{v:.10e}" for v in objs), cgs)
    if ref:
        print("   rel vs oracle", " ".join(f"{abs(a-b)/abs(b):.1e}" for a, b in zip(objs, ref["obj"])))
    p.close()
