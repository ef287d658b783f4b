"""Per-block CG iteration counts, GPU (fp64) vs oracle, on a synthetic shape (debug aid)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "one-class-ffm_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import numpy as np
import ocffm, pyoracle, synth

shape = sys.argv[1] if len(sys.argv) > 1 else "C3s"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ds = synth.generate(shape, seed=3, test_rows=300, cold_rows=5)
prm = dict(k=k, lam=4.0, omega=2.0 ** -7, r=-1.0, self_side=True, freq=False)
o = pyoracle.Oracle(ds, **prm)
p = ocffm.Problem(ds, dtype=ocffm.F32 if os.environ.get("DBG_F32") else ocffm.F64, **prm)
rng = np.random.default_rng(11)
for f1, f2 in o.blocks():
    for which in "WH":
        sc = float(os.environ.get("DBG_SCALE", 0.1 / np.sqrt(k)))
        w = rng.uniform(-sc, sc, size=(o.block_rows(f1, f2, which), k))
        o.set_block(f1, f2, which, w); p.set_block(f1, f2, which, w)
o.init_state(); p.init_state()
fu = ds.users.f
order = sorted(o.blocks(), key=lambda b: (2 if b[0] < fu <= b[1] else (0 if b[1] < fu else 1), b))
for f1, f2 in order:
    b0 = o.cg_iters_total(); p.reset_stats()
    for which in "WH":
        G = o.grad(f1, f2, which)
        Gp = p.grad(f1, f2, which)
        S, it = p.cg(f1, f2, which, G)
        So, ito = o.cg(f1, f2, which, G) if hasattr(o, "cg") else (None, None)
        print(f"  half ({f1},{f2},{which}) |G|2={float((G*G).sum()):.6e} gerr={float(np.abs(G-Gp).max()/max(1e-300,np.abs(G).max())):.2e} cg gpu={it} oracle={ito}")
    o.solve_block(f1, f2); p.solve_block(f1, f2)
    print(f"block ({f1},{f2}) cg oracle={o.cg_iters_total()-b0} gpu={int(p.stats().cg_iters)} objective oracle={o.func():.12e} gpu={p.objective():.12e}")
