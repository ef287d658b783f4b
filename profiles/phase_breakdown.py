import os, sys, time
os.environ["OCFFM_PROFILE"]="2"
sys.path.insert(0,"one-class-ffm_b200")
import ocffm, synth
ds=synth.generate("C2",seed=1)
p=ocffm.Problem(ds,k=32,lam=4.0,omega=2**-7,r=-1.0)
p.init_model(seed=1); p.init_state()
for _ in range(3): p.one_epoch()
p.reset_stats()
t=time.perf_counter()
for _ in range(5): p.one_epoch()
p.synchronize(); wall=(time.perf_counter()-t)/5*1e3
s=p.stats()
print("wall ms/epoch",wall,"cg",s.cg_iters/5,"launches",s.kernel_launches/5)
print("side  grad %.2f cg %.2f upd %.2f"%(s.ms_side_grad/5,s.ms_side_cg/5,s.ms_side_update/5))
print("cross grad %.2f cg %.2f upd %.2f"%(s.ms_cross_grad/5,s.ms_cross_cg/5,s.ms_cross_update/5))
print("hv ms/epoch",s.hv_ms/5,"launches",s.hv_launches/5)
