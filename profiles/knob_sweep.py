"""Schedule-threshold sweep on one resident shape (one GPU): for each environment variant a fresh
context, 4 warm outer iterations and 6 timed (2 and 3 on the big shapes).  python profiles/knob_sweep.py [shape] [k]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "one-class-ffm_b200"))
import ocffm, synth

shape = sys.argv[1] if len(sys.argv) > 1 else "C2"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 32
ds = synth.generate(shape, seed=1)
WARM, TIMED = (4, 6) if shape in ("C2", "C5") else (2, 3)
VARIANTS = [
    {}, {"OCFFM_MROW_MIN": "32"}, {"OCFFM_MROW_MIN": "64"}, {"OCFFM_MROW_MIN": "96"},
    {"OCFFM_MROW_ITERS": "5"}, {"OCFFM_MROW_ITERS": "12"}, {"OCFFM_MROW": "0"},
    {"OCFFM_CHUNK": "16"}, {"OCFFM_CHUNK": "64"}, {"OCFFM_HOT_MIN": "512"}, {"OCFFM_HOT_MIN": "8192"},
    {"OCFFM_PERSIST_CG": "1"}, {"OCFFM_FUSED_DOT": "0"}, {},
] if shape in ("C2", "C5") else [
    {}, {"OCFFM_MROW_MIN": "32"}, {"OCFFM_MROW_MIN": "96"}, {"OCFFM_MROW_ITERS": "5"}, {"OCFFM_CHUNK": "64"},
    {"OCFFM_HOT_MIN": "8192"}, {"OCFFM_HOT_MIN": "512"}, {"OCFFM_PERSIST_CG": "1"},
]
for env in VARIANTS:
    for key, val in env.items():
        os.environ[key] = val
    p = ocffm.Problem(ds, k=K, lam=4.0, omega=2 ** -7, r=-1.0)
    for key in env:
        del os.environ[key]
    p.init_model(seed=1)
    p.init_state()
    for _ in range(WARM):
        p.one_epoch()
    p.reset_stats()
    p.synchronize(); t = time.perf_counter()
    for _ in range(TIMED):
        p.one_epoch()
    p.synchronize(); ms = (time.perf_counter() - t) / TIMED * 1e3
    s = p.stats()
    print(f"{shape} {env or 'default'}: {ms:.2f} ms/epoch  cg {s.cg_iters}  launches {s.kernel_launches}  obj {p.objective():.6e}", flush=True)
    p.close()
