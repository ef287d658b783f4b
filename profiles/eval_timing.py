import os, sys, time
sys.path.insert(0,"one-class-ffm_b200")
import ocffm, synth, numpy as np
ds=synth.generate("C2",seed=1,test_rows=30000)
p=ocffm.Problem(ds,k=32,lam=4.0,omega=2**-7,r=-1.0)
p.init_model(seed=1); p.init_state()
p.validate(want_topk=False)
t=time.perf_counter()
for _ in range(3): p.validate(want_topk=False)
print("validate ms %.2f" % ((time.perf_counter()-t)/3*1e3), os.environ.get("OCFFM_TC_DEBUG"), os.environ.get("OCFFM_EVAL_MC"))
