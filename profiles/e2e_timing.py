"""Where an end-to-end step goes (one GPU): H2D of the model, init_state, one_epoch with and without
host mirrors, serial D2H.  python profiles/e2e_timing.py [shape] [k]"""
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
for _ in range(3):
    p.one_epoch()
host = {}
for key in model:
    rows = p.block_rows(*key)
    host[key] = p.get_block(*key, out=torch.empty((rows, K), dtype=torch.float64, pin_memory=True).numpy())
nbytes = sum(v.nbytes for v in host.values())

def timed(fn, n=3):
    p.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    p.synchronize(); return (time.perf_counter() - t) / n * 1e3

t_h2d = timed(lambda: [p.set_block(k[0], k[1], k[2], w) for k, w in host.items()])
t_init = timed(p.init_state)
t_epoch = timed(p.one_epoch)
t_d2h = timed(lambda: [p.get_block(*k, out=host[k]) for k in host])
for k, w in host.items():
    p.mirror_block(k[0], k[1], k[2], w)
t_epoch_m = timed(p.one_epoch)
def e2e_step():
    for k, w in host.items():
        p.set_block(k[0], k[1], k[2], w)
    p.init_state()
    p.one_epoch()
t_e2e = timed(e2e_step)
print(f"{shape}: e2e step (H2D + init_state + one_epoch with mirrors) {t_e2e:.2f} ms")
print(f"model {nbytes/1e6:.0f} MB  H2D {t_h2d:.2f} ms ({nbytes/t_h2d/1e6:.1f} GB/s)  init_state {t_init:.2f}  one_epoch {t_epoch:.2f}  "
      f"serial D2H {t_d2h:.2f} ms ({nbytes/t_d2h/1e6:.1f} GB/s)  one_epoch with mirrors {t_epoch_m:.2f} ms")
