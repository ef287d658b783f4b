"""Worker of tests/test_multi_gpu.py (launched by torch.distributed.run, one rank per GPU):
solves the same seeded problem sharded over WORLD_SIZE GPUs and saves what rank 0 sees."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "one-class-ffm_b200"))
import dist_util  # noqa: E402
import ocffm  # noqa: E402
import synth  # noqa: E402


def main():
    out_path, dtype = sys.argv[1], sys.argv[2]
    rank, world, local_rank = dist_util.init()
    uid = dist_util.share_unique_id(ocffm.comm_unique_id)
    comm = None if world == 1 else (world, rank, uid)
    ds = synth.generate("C1", seed=3, scale=0.3, test_rows=400, cold_rows=6)
    p = ocffm.Problem(ds, k=16, lam=4.0, omega=2.0 ** -7, r=-1.0, device=local_rank, comm=comm,
                      dtype=ocffm.F64 if dtype == "f64" else ocffm.F32)
    p.init_model(seed=5)
    p.init_state()
    objs, cgs = [], []
    for _ in range(2):
        p.reset_stats()
        p.one_epoch()
        cgs.append(int(p.stats().cg_iters))
        objs.append(p.objective())
    res = p.validate(want_topk=False)
    if rank == 0:
        blocks = {f"W{f1}_{f2}": p.get_block(f1, f2, "W") for f1, f2 in p.blocks()}
        np.savez(out_path, objs=np.array(objs), cgs=np.array(cgs), a=p.vec("a"), b=p.vec("b"),
                 prec=res["prec"], ndcg=res["ndcg"], ploss=res["ploss"], **blocks)
    dist_util.barrier()


if __name__ == "__main__":
    main()
