#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: sec per outer iteration and nnz/s at d=32
on the KKBox-shaped synthetic set (BASELINE.json configs[1], SURVEY.md 8 shape C2), plus
full-ranking eval users/s, measured through the C ABI of libocffm_cuda.so.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C1..C5]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A step is one outer iteration (one_epoch, ffm.cpp:852-870: every side + cross block solve and
cache_sasb) over the whole synthetic set.  `value` = nnz/s = N_trav / time with N_trav as defined
in SURVEY.md 8(d) (it normalises out CG-count differences), inputs resident in HBM; the library's
counters are per rank and summed over the ranks.  `e2e` = the same metric when every step starts
from a HOST-resident fp64 model and ends with the host copy current again: H2D of all W/H from
pinned memory (ocffm_set_block), ocffm_init_state, ocffm_one_epoch with every block registered as
a host mirror (ocffm_mirror_block: D2H of each block on a second stream right after its solve).
`roofline`: the hs_cross row pass (algorithmic bytes over the pass time; a phase of the persistent
CG kernel, stamped inside the kernel), `roofline.cg_kernel`: the fused kernel as a whole; all
fractions are per GPU.  N > 1: C2 scales WEAK (one block of 30 000 users per GPU), the other
workloads STRONG (the fixed shape sharded by rows), and every run first checks an N-rank fp64
solve of a small set against a 1-rank solve (`multi_rank_parity`).
`--impl reference` times the UNMODIFIED reference's own one_epoch()/validate()
(oracle/_ref/ref_harness_blas, built from /root/reference by oracle/Makefile) on the box's host
cores: the FULL C2 shape (1 warm-up + at most 2 timed outer iterations), stated sub-samples for
the multi-million-row shapes.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "one-class-ffm_b200"))

HYPER = dict(lam=4.0, omega=2.0 ** -7, r=-1.0)     # script/grid.sh:237 canonical flags
WORKLOADS = {
    # name -> (shape, k, test rows)
    "C2": ("C2", 32, 30_000),
    "C1": ("C1", 16, 2_000),
    "C3": ("C3", 16, 20_000),
    "C4": ("C4", 32, 20_000),
    "C5": ("C2", 64, 30_000),   # BASELINE configs[4]: d=64 KKBox-shaped scoring + top-k sweep (see "eval")
}
SHAPE_M = {"C1": 10_000, "C2": 30_000, "C3": 8_000_000, "C4": 4_000_000, "C5": 30_000}
CPU_SAMPLE_SCALE = {"C2": 0.2, "C1": 1.0, "C3": 0.02, "C4": 0.02, "C5": 0.1}      # cpu_baseline leg of the GPU arm
# --impl reference: the headline workload runs at FULL size (same config as the GPU arm, fewer timed
# iterations); the multi-million-row shapes stay on a stated sub-sample
REF_ARM_SCALE = {"C2": 1.0, "C1": 1.0, "C3": 0.02, "C4": 0.02, "C5": 0.1}
REF_ARM_MAX_STEPS = 2
# BASELINE configs[2..4] shard ONE fixed set over the GPUs (strong scaling); the headline C2 run keeps
# round 1's weak scaling (one block of 30 000 users per GPU over the same items)
DEFAULT_SCALING = {"C2": "weak", "C1": "strong", "C3": "strong", "C4": "strong", "C5": "strong"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d.get("bf16_tflops", 1590.0)),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.sm_max = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=self.sm_max, reasons=sorted(self.reasons))
        return dict(sm_mhz=float(np.median(self.samples)), sm_max_mhz=self.sm_max,
                    reasons=sorted(self.reasons), samples=len(self.samples))


def run_port_arm(args, shape, k, test_rows):
    """oracle/_ref was not built (no /root/reference on this machine): times the plain-C restatement
    of the same path (oracle/liboracle.so, one thread) on a smaller sample.  kind = "port"."""
    import synth
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import pyoracle
    except Exception:
        return None
    scale = (args.cpu_scale or CPU_SAMPLE_SCALE[args.workload]) * 0.25
    ds = synth.generate(shape, seed=args.seed, scale=scale, test_rows=max(64, int(test_rows * scale * 0.1)))
    o = pyoracle.Oracle(ds, k=k, **HYPER)
    o.init_model_rng()
    o.init_state()
    nnz_y = int(ds.train.idx.size)
    fu = ds.users.f
    nnzx = [int(f.idx.size) for f in ds.users.fields] + [int(f.idx.size) for f in ds.items.fields]
    n_trav, sec, cgs = 0.0, 0.0, []
    t_all = time.time()
    for e in range(args.warmup + args.steps):
        timed = e >= args.warmup
        before_epoch = o.cg_iters_total()
        t0 = time.time()
        order = sorted(o.blocks(), key=lambda b: (2 if b[0] < fu <= b[1] else (0 if b[1] < fu else 1), b))
        for f1, f2 in order:               # one_epoch order: user side, item side, cross (ffm.cpp:852-870)
            before = o.cg_iters_total()
            o.solve_block(f1, f2)
            c = o.cg_iters_total() - before
            if timed:
                xa, xb, cross = nnzx[f1], nnzx[f2], f1 < fu <= f2
                n_trav += 2 * nnz_y + xa + xb + c * ((nnz_y if cross else 0) + (xa + xb) / 2.0) \
                    + 2 * (2 * nnz_y) + xa + xb
        if timed:
            sec += time.time() - t0
            cgs.append(o.cg_iters_total() - before_epoch)
    t0 = time.time()
    o.validate(want_topk=False)
    t_val = time.time() - t0
    sample = (f"{shape} scaled x{scale} (m={ds.m}, n={ds.n}, nnz_y={nnz_y}), k={k}, {args.steps} timed epochs "
              f"after {args.warmup} warm-up, plain-C port of the reference path, 1 thread")
    return dict(kind="port", value=n_trav / sec, sec_per_epoch=sec / max(1, args.steps), cores=1,
                host_cores=os.cpu_count() or 1, sample=sample, cg_iters=cgs, wall_s=time.time() - t_all,
                eval_users_per_s=ds.test.rows / t_val if ds.test is not None and t_val > 0 else None,
                harness="liboracle.so")


def run_reference_arm(args, shape, k, test_rows):
    """Times the unmodified reference (oracle/_ref) on host cores on a bounded sample."""
    import synth
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness_blas")
    kind = "reference"
    if not os.path.exists(harness):
        harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    if not os.path.exists(harness):
        return run_port_arm(args, shape, k, test_rows)
    scale = args.cpu_scale or (REF_ARM_SCALE if args.impl == "reference" else CPU_SAMPLE_SCALE)[args.workload]
    # the evaluator is timed on a bounded number of test rows (the reference ranks ~50-900 users/s)
    ds = synth.generate(shape, seed=args.seed, scale=scale, test_rows=min(1000, max(64, int(test_rows * scale * 0.1))))
    ncpu = os.cpu_count() or 1
    threads = args.cpu_threads or min(ncpu, 16)
    epochs = args.warmup + args.steps
    with tempfile.TemporaryDirectory() as tmp:
        item_p, tr_p, te_p = synth.write_text(ds, tmp)
        env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS=str(threads))
        cmd = [harness, "time", item_p, tr_p, te_p, str(epochs), "-k", str(k), "-l", str(HYPER["lam"]),
               "-w", str(HYPER["omega"]), "-r", str(HYPER["r"]), "-c", str(threads)]
        t0 = time.time()
        out = subprocess.check_output(cmd, env=env, text=True)
        wall = time.time() - t0
    res = json.loads(out.strip().splitlines()[-1])
    ep = res["epochs"][args.warmup:]
    nnz_y = int(ds.train.idx.size)
    fu, fv = ds.users.f, ds.items.f
    nnzx_u = [int(f.idx.size) for f in ds.users.fields]
    nnzx_v = [int(f.idx.size) for f in ds.items.fields]
    # per-block CG counts in one_epoch order (harness prints them); N_trav per SURVEY.md 8(d) with
    # a block's CG iterations split evenly between its two halves for the (small) nnzX term
    n_trav = 0.0
    for e in ep:
        for blk in e["blocks"]:
            f1, f2, c = blk["f1"], blk["f2"], blk["cg"]
            xa = nnzx_u[f1] if f1 < fu else nnzx_v[f1 - fu]
            xb = nnzx_u[f2] if f2 < fu else nnzx_v[f2 - fu]
            cross = f1 < fu <= f2
            n_trav += 2 * nnz_y + xa + xb                       # two gradients
            n_trav += c * ((nnz_y if cross else 0) + (xa + xb) / 2.0)
            n_trav += 2 * (2 * nnz_y) + xa + xb                 # two updates
    sec = sum(e["sec"] for e in ep)
    value = n_trav / sec
    sample = (f"{shape} scaled x{scale} (m={ds.m}, n={ds.n}, nnz_y={nnz_y}), k={k}, {len(ep)} timed epochs "
              f"after {args.warmup} warm-up, {threads} OpenMP threads, OpenBLAS pinned to 1 thread")
    return dict(kind=kind, value=value, sec_per_epoch=sec / len(ep), cores=threads, host_cores=ncpu,
                shape_keys=dict(m=ds.m, n=ds.n, nnz_y=nnz_y, fu=fu, fv=fv, scale=scale,
                                eval_rows_timed=0 if ds.test is None else ds.test.rows),
                sample=sample, cg_iters=[e["cg_iters"] for e in ep], wall_s=wall,
                eval_users_per_s=(res["m_t"] / res["validate_s"]) if "validate_s" in res else None,
                harness=os.path.basename(harness))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--ns", action="store_true", help="--ns: cross blocks only")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debug only)")
    ap.add_argument("--cpu-scale", type=float, default=0.0)
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eval", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="N > 1: weak = one block of the shape's users per GPU, strong = the fixed shape sharded")
    ap.add_argument("--no-parity-check", action="store_true", help="N > 1: skip the 1-rank vs N-rank fp64 self-check")
    args = ap.parse_args()
    if args.impl == "ours":
        args.warmup = max(args.warmup, 3)
    else:   # bounded: the whole reference run must end within a few minutes
        args.warmup, args.steps = 1, max(1, min(args.steps, REF_ARM_MAX_STEPS))
    scaling = args.scaling or DEFAULT_SCALING[args.workload]

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    shape, k, test_rows = WORKLOADS[args.workload]
    config = dict(workload=f"{args.workload}: KKBox-shaped synthetic one-class set" if shape == "C2"
                  else f"{args.workload} synthetic one-class set",
                  shape=shape, k=k, lam=HYPER["lam"], omega=HYPER["omega"], r=HYPER["r"],
                  self_side=not args.ns, seed=args.seed, zipf=1.3, l2="inputs larger than L2 (126 MB)")

    if args.impl == "reference":
        if rank != 0:
            return
        ref = run_reference_arm(args, shape, k, test_rows)
        if ref is None:
            print(json.dumps(dict(impl="reference", unavailable="oracle/_ref binaries missing (make -C oracle ref)")))
            return
        config.update(sample=ref["sample"], **ref.get("shape_keys", {}))
        line = dict(metric="nnz_per_s", value=ref["value"], unit="nnz/s", n_gpus=0, steps=args.steps,
                    warmup=args.warmup, ms_per_step=1e3 * ref["sec_per_epoch"], higher_is_better=True,
                    scaling=scaling, vs_baseline=None, dtype="f64", data="synthetic", config=config,
                    impl="reference", sec_per_outer_iteration=ref["sec_per_epoch"],
                    eval_users_per_s=ref["eval_users_per_s"], gpu_launches=0,
                    cpu_baseline=dict(value=ref["value"], unit="nnz/s", cores=ref["cores"], kind=ref["kind"],
                                      sample=ref["sample"]),
                    e2e=dict(value=ref["value"], unit="nnz/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import dist_util
    import ocffm
    import synth

    if ocffm.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; libocffm_cuda has no CPU fallback")
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        dist_util.init("nccl")
        comm = (world, rank, dist_util.share_unique_id(ocffm.comm_unique_id))

    t_gen = time.time()
    n_test = 0 if args.no_eval else max(64, int(test_rows * args.scale))
    if world > 1 and scaling == "weak":
        # WEAK scaling: every GPU brings one block of the shape's users (m = 30k x N for C2) over
        # the same items; each rank generates its own seeded block, blocks are exchanged through
        # /tmp (one node), every rank assembles the full set (it needs both orientations of Omega)
        import pickle
        tag = os.path.join(tempfile.gettempdir(), f"ocffm_{shape}_{args.seed}_w{world}")
        blk = synth.user_block(shape, args.seed, rank, world)
        with open(f"{tag}_b{rank}.tmp", "wb") as fh:
            pickle.dump(blk, fh, protocol=4)
        os.replace(f"{tag}_b{rank}.tmp", f"{tag}_b{rank}.pkl")
        dist.barrier()
        blocks = []
        for b in range(world):
            with open(f"{tag}_b{b}.pkl", "rb") as fh:
                blocks.append(pickle.load(fh))
        ds = synth.assemble_blocks(shape, args.seed, blocks, test_rows=n_test)
        dist.barrier()
        if rank == 0:
            for b in range(world):
                os.remove(f"{tag}_b{b}.pkl")
    else:
        # one GPU, or STRONG scaling: every rank builds the same seeded set and keeps its row slices
        ds = synth.generate(shape=shape, seed=args.seed, scale=args.scale, test_rows=n_test)
    t_gen = time.time() - t_gen
    config.update(m=ds.m, n=ds.n, nnz_y=int(ds.train.idx.size), fu=ds.users.f, fv=ds.items.f,
                  m_t=0 if ds.test is None else ds.test.rows, parallelism=f"rows sharded over {world} GPU(s)" + (
                      "" if world == 1 else (f"; weak scaling: {world} blocks of {SHAPE_M.get(shape, 0)} users over the same items"
                                             if scaling == "weak" else "; strong scaling: the fixed shape sharded by rows")))
    dtype = ocffm.F32 if args.dtype == "f32" else ocffm.F64
    os.environ.setdefault("OCFFM_PROFILE", "1")     # CUDA events around every hv_cross launch
    prob = ocffm.Problem(ds, k=k, dtype=dtype, device=local_rank, self_side=not args.ns, comm=comm, **HYPER)
    # the multi-million-row shapes draw their model on the GPU (ocffm_init_model: no 10 GB host copy per
    # rank); the others with numpy on the host, as round 1 did
    big_model = shape in ("C3", "C4")
    if big_model:
        prob.init_model_device(seed=args.seed)
        model = {(f1, f2, w): None for f1, f2 in prob.blocks() for w in "WH"}
    else:
        model = prob.init_model(seed=args.seed)
    prob.init_state()
    stream = torch.cuda.ExternalStream(prob.stream(), device=torch.device("cuda", local_rank))

    def barrier():
        prob.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        prob.one_epoch()
    barrier()
    prob.reset_stats()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        prob.one_epoch()
    ev1.record(stream)
    barrier()
    sampler.stop_flag = True
    ms = ev0.elapsed_time(ev1)
    st = prob.stats()
    ms = dist_util.max_over_ranks(ms)
    sec = ms / 1e3
    # the library's counters are per rank (this rank's rows): the job's totals are sums over ranks
    nnz_total, cg_total_max, algo_total = dist_util.sum_over_ranks(float(st.nnz_traversed)), st.cg_iters, \
        dist_util.sum_over_ranks(float(st.algo_bytes))
    value = nnz_total / sec
    objective = prob.objective()

    pk = peaks()
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r02_hv_traffic.json")
    if args.workload == "C2" and world == 1 and os.path.exists(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh)["traffic_bytes_per_launch_avg"]   # ncu dram read+write per launch
        traffic_src = "profiles/r02_hv_traffic.json (one ncu --set full capture of the same command, NOT measured in this run)"
    # PER-GPU roofline: this rank's algorithmic bytes over this rank's kernel time against ONE GPU's peak
    hv_gbs = (st.hv_algo_bytes / 1e9) / (st.hv_ms / 1e3) if st.hv_ms > 0 else None
    roofline = dict(bound="hbm", kernel="hs_cross row pass: k_hess_heavy (per-row Gram blocks streamed) + k_hess_cross "
                    "(gathers, light rows)" if st.row_gram_builds else "k_hess_cross (hs_cross row pass)",
                    achieved=hv_gbs, peak=pk["hbm_gbs"],
                    unit="GB/s", frac=(hv_gbs / pk["hbm_gbs"]) if hv_gbs else None, traffic=traffic,
                    traffic_source=traffic_src, per_gpu=True,
                    peak_source=pk["source"], launches=int(st.hv_launches),
                    avg_launch_ms=(st.hv_ms / st.hv_launches) if st.hv_launches else None,
                    share_of_step=(st.hv_ms / ms) if ms > 0 else None,
                    algo_bytes_per_launch=(st.hv_algo_bytes / st.hv_launches) if st.hv_launches else None,
                    whole_epoch_gbs_per_gpu=(st.algo_bytes / 1e9) / sec,
                    whole_epoch_frac_per_gpu=(st.algo_bytes / 1e9) / sec / pk["hbm_gbs"],
                    whole_epoch_gbs_all_gpus=(algo_total / 1e9) / sec,
                    row_gram_builds=int(st.row_gram_builds), row_gram_bytes=int(st.row_gram_bytes))
    if st.cg_kernel_launches:
        # Cross halves solved by the persistent CG kernel: the hs_cross row pass is a PHASE of that kernel.
        # Its time is stamped inside the kernel (%globaltimer of CTA 0 between the grid-wide barriers that
        # bracket the phase, OCFFM_PROFILE=1) and added to the event-timed separate launches above, so
        # `achieved` / `frac` keep round 1's definition (hs_cross algorithmic bytes over the pass time).
        # `cg_kernel` is the fused kernel as a whole: ALL its algorithmic bytes (per iteration the hs_cross
        # formula + 7 D k s of CG vector traffic, SURVEY 8(a) A7 + A8) over its event-timed duration.
        ck_gbs = (st.cg_kernel_algo_bytes / 1e9) / (st.cg_kernel_ms / 1e3)
        roofline.update(
            kernel="hs_cross row pass: phase of k_cg_cross_persist (one cooperative launch per cross half solve) "
                   "+ k_hess_cross / k_hess_heavy launches of the halves on the per-iteration kernels",
            timing="%globaltimer stamps between the grid barriers for the phases, CUDA events for the launches",
            cg_kernel=dict(kernel="k_cg_cross_persist (per CG iteration: direction + V*QTQ, hs_cross row pass, step)",
                           achieved=ck_gbs, frac=ck_gbs / pk["hbm_gbs"], launches=int(st.cg_kernel_launches),
                           avg_launch_ms=st.cg_kernel_ms / st.cg_kernel_launches,
                           avg_cg_iteration_ms=st.cg_kernel_ms / max(1, st.cg_kernel_iters),
                           cg_iterations=int(st.cg_kernel_iters),
                           share_of_step=st.cg_kernel_ms / ms if ms > 0 else None,
                           algo_bytes_per_launch=st.cg_kernel_algo_bytes / st.cg_kernel_launches))
    free_b, total_b = torch.cuda.mem_get_info(local_rank)
    footprint = dict(omega_device_bytes=int(st.omega_device_bytes), hbm_used_bytes=int(total_b - free_b),
                     note="per rank (rank 0): Omega slices = row pointers, column ids, y-tilde and work-item lists of both orientations")

    # ---- evaluation (validate, ffm.cpp:925-1016) --------------------------------------------
    eval_info = None
    if ds.test is not None:
        prob.validate(want_topk=False)          # warm-up
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        res = prob.validate(want_topk=False)
        e1.record(stream)
        barrier()
        ems = e0.elapsed_time(e1)
        ems = dist_util.max_over_ranks(ems)
        fx = ds.users.f * ds.items.f
        flops = 2.0 * ds.test.rows * ds.train.n_items * fx * k
        # tensor roofline of the scorer: 3xTF32 issues 3 MMAs per algorithmic MAC; TF32 dense peak is
        # taken as half of the measured bf16 burst figure
        tf32_peak = pk["bf16_tflops"] / 2.0
        alg_tflops = flops / (ems / 1e3) / 1e12
        eval_info = dict(users_per_s=ds.test.rows / (ems / 1e3), ms=ems, m_t=ds.test.rows,
                         n_ranked=ds.train.n_items, tflops=alg_tflops,
                         roofline=dict(bound="tensor", kernel="k_score_topk_tc (tcgen05 kind::tf32, 3xTF32)"
                                       if args.dtype == "f32" else "k_score_topk (SIMT)",
                                       achieved=(3.0 * alg_tflops if args.dtype == "f32" else alg_tflops) / world,
                                       peak=tf32_peak, unit="TFLOP/s", per_gpu=True,
                                       frac=(3.0 * alg_tflops if args.dtype == "f32" else alg_tflops) / world / tf32_peak,
                                       useful_frac=alg_tflops / world / tf32_peak,
                                       note="whole validate() timed (SpMM, TF32 split, scorer, merge, metrics), per GPU; "
                                            "achieved counts the 3 TF32 MMAs per algorithmic MAC (useful_frac counts one); ncu's own "
                                            "tensor-pipe-active for the kernel alone is in profiles/"),
                         p_at_10=float(res["prec"][1]), ndcg_at_10=float(res["ndcg"][1]), ploss=float(res["ploss"]))

    # ---- end to end through the C ABI with host buffers --------------------------------------
    e2e_steps = min(args.steps, 5)
    bytes_model = int(sum(prob.block_rows(*key) * k * 8 for key in model))
    # every rank keeps a pinned fp64 copy of the whole model for this leg: skipped when that would pin
    # more than 32 GB of host memory on the node (C4 at N = 8: 8 x 12 GB)
    if args.no_e2e or bytes_model * world > (32 << 30):
        e2e = dict(value=None, unit="nnz/s", h2d_bytes_per_step=bytes_model, d2h_bytes_per_step=bytes_model, steps=0,
                   what=f"skipped: {bytes_model / 2**30:.1f} GiB of fp64 model per rank x {world} ranks of pinned host memory")
    else:
        # the model lives in PINNED host memory (the library DMAs straight from / into it)
        host_model = {}
        for key in model:
            rows = prob.block_rows(*key)
            buf = torch.empty((rows, k), dtype=torch.float64, pin_memory=True).numpy()
            host_model[key] = prob.get_block(*key, out=buf)
        for key, w in host_model.items():
            prob.mirror_block(key[0], key[1], key[2], w)   # one_epoch() keeps these host copies current

        def e2e_step():
            for key, w in host_model.items():
                prob.set_block(key[0], key[1], key[2], w)      # H2D of the whole model
            prob.init_state()
            prob.one_epoch()                                   # D2H of every block inside (second stream)

        e2e_step()       # one untimed warm-up: the first mirrored epoch allocates the 8 staging buffers
        prob.synchronize()   # (0.7 GB of cudaMalloc, 150-340 ms once: profiles/r02_e2e_steps.txt)
        barrier()
        prob.reset_stats()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        prob.synchronize()
        e2e_sec = time.perf_counter() - t0
        e2e_sec = dist_util.max_over_ranks(e2e_sec)
        st2 = prob.stats()
        e2e = dict(value=dist_util.sum_over_ranks(float(st2.nnz_traversed)) / e2e_sec, unit="nnz/s",
                   h2d_bytes_per_step=bytes_model, d2h_bytes_per_step=bytes_model, steps=e2e_steps, warmup=1,
                   sec_per_step=e2e_sec / e2e_steps,
                   what="per step: ocffm_set_block for every W/H (H2D from pinned fp64 host arrays), ocffm_init_state, "
                        "ocffm_one_epoch with every W/H registered as a host mirror (ocffm_mirror_block): each block is "
                        "copied back (D2H, fp64) into the same pinned arrays right after its solve, on a second stream")
        for key in host_model:
            prob.mirror_block(key[0], key[1], key[2], None)

    # ---- N > 1: fp64 self-check, N ranks against ONE rank on the same small set ------------------
    parity = None
    if world > 1 and not args.no_parity_check:
        small = synth.generate("C1", seed=args.seed + 7, scale=0.2, test_rows=64)     # identical on every rank
        sprm = dict(k=16, self_side=True, **HYPER)
        pn = ocffm.Problem(small, dtype=ocffm.F64, device=local_rank,
                           comm=(world, rank, dist_util.share_unique_id(ocffm.comm_unique_id)), **sprm)
        pn.init_model(seed=3)
        pn.init_state()
        pn.reset_stats()
        pn.one_epoch()
        pn.one_epoch()
        obj_n, cg_n, val_n = pn.objective(), int(pn.stats().cg_iters), pn.validate(want_topk=False)
        pn.close()
        if rank == 0:
            p1 = ocffm.Problem(small, dtype=ocffm.F64, device=local_rank, **sprm)
            p1.init_model(seed=3)
            p1.init_state()
            p1.reset_stats()
            p1.one_epoch()
            p1.one_epoch()
            obj_1, cg_1, val_1 = p1.objective(), int(p1.stats().cg_iters), p1.validate(want_topk=False)
            p1.close()
            parity = dict(what="C1 x0.2, k=16, fp64, 2 outer iterations + validate: N-rank context vs 1-rank context",
                          objective_1rank=obj_1, objective_nrank=obj_n, cg_1rank=cg_1, cg_nrank=cg_n,
                          ndcg10_1rank=float(val_1["ndcg"][1]), ndcg10_nrank=float(val_n["ndcg"][1]),
                          ok=bool(cg_1 == cg_n and abs(obj_1 - obj_n) <= 1e-9 * abs(obj_1)
                                  and abs(val_1["ndcg"][1] - val_n["ndcg"][1]) <= 1e-9))
            if not parity["ok"]:
                print("bench.py: multi-rank parity self-check FAILED: " + json.dumps(parity), file=sys.stderr)
        dist.barrier()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            ref_args = argparse.Namespace(**vars(args))
            ref_args.steps, ref_args.warmup = 2, 1
            r = run_reference_arm(ref_args, shape, k, test_rows)
            if r:
                cpu = dict(value=r["value"], unit="nnz/s", cores=r["cores"], kind=r["kind"], sample=r["sample"],
                           sec_per_epoch=r["sec_per_epoch"], eval_users_per_s=r["eval_users_per_s"],
                           host_cores=r["host_cores"])
        except Exception as exc:  # the baseline is reported, never required
            cpu = dict(value=None, unit="nnz/s", cores=0, kind="reference", sample=f"failed: {exc}")

    if rank == 0:
        line = dict(metric="nnz_per_s", value=value, unit="nnz/s", n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True,
                    scaling=scaling, vs_baseline=None, dtype=args.dtype, data="synthetic", config=config,
                    sec_per_outer_iteration=sec / args.steps, cg_iters_per_step=st.cg_iters / args.steps,
                    objective=objective, gpu_launches=int(st.kernel_launches), e2e=e2e, roofline=roofline,
                    cpu_baseline=cpu, eval=eval_info, clocks=sampler.summary(), datagen_s=t_gen,
                    footprint=footprint, multi_rank_parity=parity)
        if int(os.environ.get("OCFFM_PROFILE", "1")) >= 2:      # diagnostic runs only (event timers per phase)
            line["phases_ms_per_step"] = {f: getattr(st, "ms_" + f) / args.steps for f in (
                "side_grad", "side_cg", "side_update", "cross_grad", "cross_cg", "cross_update")}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
