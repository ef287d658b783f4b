"""One warm outer iteration on a synthetic shape, for `ncu` launch lists:
python profiles/one_epoch.py [shape] [k] [warm epochs]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "one-class-ffm_b200"))
import ocffm  # noqa: E402
import synth  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "C2"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 32
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ds = synth.generate(shape, seed=1)
p = ocffm.Problem(ds, k=k, lam=4.0, omega=2 ** -7, r=-1.0)
p.init_model(seed=1)
p.init_state()
for _ in range(warm):
    p.one_epoch()
p.reset_stats()
p.one_epoch()
p.synchronize()
s = p.stats()
print(shape, "launches", s.kernel_launches, "cg", s.cg_iters)
