"""Warm per-phase breakdown of one outer iteration (OCFFM_PROFILE=2 event timers inside the
library): python profiles/phase_breakdown.py [shape] [k]"""
import os
import sys
import time

os.environ["OCFFM_PROFILE"] = "2"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "one-class-ffm_b200"))
import ocffm  # noqa: E402
import synth  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "C2"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 32
ds = synth.generate(shape, seed=1)
p = ocffm.Problem(ds, k=k, lam=4.0, omega=2 ** -7, r=-1.0)
p.init_model(seed=1)
p.init_state()
for _ in range(3):
    p.one_epoch()
p.reset_stats()
n = 5 if shape == "C2" else 2
t = time.perf_counter()
for _ in range(n):
    p.one_epoch()
p.synchronize()
wall = (time.perf_counter() - t) / n * 1e3
s = p.stats()
print(shape, "wall ms/epoch %.2f" % wall, "cg", s.cg_iters / n, "launches", s.kernel_launches / n)
print("side  grad %.2f cg %.2f upd %.2f" % (s.ms_side_grad / n, s.ms_side_cg / n, s.ms_side_update / n))
print("cross grad %.2f cg %.2f upd %.2f" % (s.ms_cross_grad / n, s.ms_cross_cg / n, s.ms_cross_update / n))
print("hv ms/epoch %.2f" % (s.hv_ms / n), "launches", s.hv_launches / n)
