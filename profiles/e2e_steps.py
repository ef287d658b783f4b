"""Per-step times of bench.py's end-to-end leg (H2D of the model, init_state, one_epoch with host
mirrors) after 8 and after 25 outer iterations.  python profiles/e2e_steps.py [shape] [k]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "one-class-ffm_b200"))
import numpy as np, torch
import ocffm, synth

shape = sys.argv[1] if len(sys.argv) > 1 else "C2"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 32
ds = synth.generate(shape, seed=1)
p = ocffm.Problem(ds, k=K, lam=4.0, omega=2 ** -7, r=-1.0)
model = p.init_model(seed=1)
p.init_state()

def now():
    p.synchronize(); return time.perf_counter()

def leg(tag):
    host = {}
    for key in model:
        rows = p.block_rows(*key)
        host[key] = p.get_block(*key, out=torch.empty((rows, K), dtype=torch.float64, pin_memory=True).numpy())
    nbytes = sum(v.nbytes for v in host.values())
    t0 = now()
    for key, w in host.items(): p.set_block(key[0], key[1], key[2], w)
    t1 = now()
    for key, w in host.items(): p.get_block(*key, out=w)
    t2 = now()
    print(f"{tag}: model {nbytes/1e6:.0f} MB  H2D {1e3*(t1-t0):.1f} ms ({nbytes/(t1-t0)/1e9:.1f} GB/s)  D2H {1e3*(t2-t1):.1f} ms ({nbytes/(t2-t1)/1e9:.1f} GB/s)")
    for key, w in host.items(): p.mirror_block(key[0], key[1], key[2], w)
    for i in range(5):
        p.reset_stats()
        a = now()
        for key, w in host.items(): p.set_block(key[0], key[1], key[2], w)
        b = now()
        p.init_state()
        c = now()
        p.one_epoch()
        d = now()
        s = p.stats()
        print(f"{tag} e2e step {i}: total {1e3*(d-a):.1f} ms = H2D {1e3*(b-a):.1f} + init_state {1e3*(c-b):.1f} + one_epoch+mirrors {1e3*(d-c):.1f}  cg {s.cg_iters} launches {s.kernel_launches} row_gram_builds {s.row_gram_builds}", flush=True)
    for key in host: p.mirror_block(key[0], key[1], key[2], None)
    a = now(); p.one_epoch(); d = now()
    print(f"{tag} plain one_epoch {1e3*(d-a):.1f} ms", flush=True)

def epochs(n):
    a = now()
    for _ in range(n): p.one_epoch()
    d = now()
    print(f"{n} outer iterations: {1e3*(d-a)/n:.2f} ms each", flush=True)

epochs(8); leg("after 8")
epochs(11); leg("after 25")
