"""Markdown tables of DESIGN.md section 6 from the committed bench lines profiles/r02_n{N}_{W}.json."""
import json, os, sys
here = os.path.dirname(os.path.abspath(__file__))
names = {"C2": "C2 KKBox-shaped, k=32 (WEAK: one block of 30 000 users per GPU over the same 360 000 items)",
         "C4": "C4 KDD12-shaped, k=32, 4 M rows x 50 k items, 8 cross pairs (STRONG: fixed set sharded)",
         "C3": "C3 Outbrain-shaped, k=16, 8 M rows x 20 k items (STRONG)",
         "C5": "C5 KKBox-shaped at k=64 (STRONG; BASELINE configs[4]: the eval sweep)"}
out = []
for w in ("C2", "C4", "C3", "C5"):
    rows = []
    for n in (1, 2, 4, 8):
        p = os.path.join(here, f"r02_n{n}_{w}.json")
        if not os.path.exists(p):
            continue
        txt = open(p).read().strip().splitlines()
        if not txt:
            continue
        d = json.loads(txt[-1])
        r, ev = d["roofline"], d.get("eval") or {}
        er = ev.get("roofline") or {}
        sep = r.get("separate_hessian_passes") or {}
        if "cg_kernel" in r or not r.get("cg_iterations"):
            hv_frac = r["frac"]                       # all Hessian passes (phases of the persistent kernel + launches)
        else:                                         # lines written before the phase stamps existed
            hv_frac = (sep.get("gbs") or 0) / r["peak"] if sep.get("gbs") else None
        rows.append((n, d["config"].get("m"), d["ms_per_step"], d["value"], d["cg_iters_per_step"], r["whole_epoch_frac_per_gpu"],
                     hv_frac, ev.get("ms"), ev.get("users_per_s"), er.get("frac"), er.get("useful_frac"),
                     (d["e2e"] or {}).get("sec_per_step"), d["footprint"]["omega_device_bytes"] / 1e6,
                     (d.get("multi_rank_parity") or {}).get("ok")))
    if not rows:
        continue
    out.append(f"**{names[w]}**\n")
    out.append("| GPUs | users | ms / outer iteration | nnz/s | speed-up | CG it. | whole-iteration HBM frac / GPU | hs_cross row pass, HBM frac / GPU | validate ms | users/s | scorer tensor frac / GPU (3 MMAs per MAC; useful) | e2e ms / step | Ω MB / rank | N-rank == 1-rank |")
    out.append("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    base = rows[0]
    for (n, m, ms, v, cg, wf, hf, ems, ups, ef, uf, e2e, om, ok) in rows:
        sp = v / base[3] if w == "C2" else base[2] / ms
        out.append(f"| {n} | {m:,} | {ms:.1f} | {v:.3g} | {sp:.2f}x | {cg:.0f} | {wf:.3f} | " + (f"{hf:.3f}" if hf else "—") +
                   f" | " + (f"{ems:.1f}" if ems else "—") + " | " + (f"{ups:.3g}" if ups else "—") + " | " +
                   (f"{ef:.3f}; {uf:.3f}" if ef is not None and uf is not None else (f"{ef:.3f}" if ef else "—")) + " | " +
                   (f"{1e3 * e2e:.1f}" if e2e else "skipped") + f" | {om:.0f} | " + ("—" if ok is None else ("yes" if ok else "NO")) + " |")
    out.append("")
print("\n".join(out))
