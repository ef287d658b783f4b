#!/usr/bin/env python
"""Mint the golden fixtures under tests/golden/ from the UNMODIFIED reference.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference and the
binaries `make -C oracle ref` puts in oracle/_ref/); the GPU box only ever reads the committed
outputs.  For every case it

  1. generates a seeded synthetic set (one-class-ffm_b200/synth.py) and writes it in the
     reference's text format (kept under tests/golden/<case>/ so the host reader can be
     checked against the reference's reader),
  2. runs oracle/_ref/ref_harness dump ... (reference reader, init, per-phase probes, epochs,
     validate) and stores the dump as a compressed .npz,
and it runs the reference's own nDCG known-answer fixture (script/nDCG_degub_tool) through
the -DEBUG_nDCG -DSHOW_SCORE_ONLY build, storing labels + per-user nDCG@10.
"""
import importlib.util
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"

sys.path.insert(0, HERE)
import pyoracle  # noqa: E402

spec = importlib.util.spec_from_file_location("ocffm_synth", os.path.join(ROOT, "one-class-ffm_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
sys.modules["ocffm_synth"] = synth
spec.loader.exec_module(synth)

CASES = [
    # name, generator kwargs, harness flags
    dict(name="tiny", gen=dict(shape="tiny", seed=7, test_rows=24, cold_rows=3),
         flags=["-k", "8", "-l", "0.5", "-w", "0.0625", "-r", "-1", "-t", "3"]),
    dict(name="tiny_ns", gen=dict(shape="tiny", seed=7, test_rows=24, cold_rows=3),
         flags=["-k", "8", "-l", "0.5", "-w", "0.0625", "-r", "-1", "-t", "3", "--ns"]),
    dict(name="tiny_freq", gen=dict(shape="tiny", seed=7, test_rows=24, cold_rows=3),
         flags=["-k", "4", "-l", "0.05", "-w", "0.125", "-r", "-0.5", "-t", "2", "--freq"]),
    dict(name="small", gen=dict(shape="C1", seed=11, scale=0.03, test_rows=48, cold_rows=4),
         flags=["-k", "16", "-l", "4", "-w", "0.0078125", "-r", "-1", "-t", "3"]),
]


def run_case(case):
    ds = synth.generate(**case["gen"])
    ds.name = case["name"]
    out_dir = os.path.join(GOLD, case["name"])
    item_p, tr_p, te_p = synth.write_text(ds, out_dir)
    with tempfile.TemporaryDirectory() as tmp:
        dump = os.path.join(tmp, "dump.ocfd")
        cmd = [os.path.join(HERE, "_ref", "ref_harness"), "dump", item_p, tr_p, te_p or "-", dump] + case["flags"]
        subprocess.check_call(cmd)
        d = pyoracle.load_ocfd(dump)
    d["meta.flags"] = np.frombuffer(" ".join(case["flags"]).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(GOLD, case["name"] + ".npz"), **d)
    print(f"{case['name']}: {len(d)} arrays, m={ds.m} n={ds.n} nnz_y={ds.train.idx.size}",
          "func", d.get("epochs.func"), "cg", d["epochs.cg_iters"])


def run_cli(case):
    """stdout and text model of the unmodified reference CLI (train.cpp) -- the drop-in contract."""
    out_dir = os.path.join(GOLD, case["name"])
    item_p, tr_p, te_p = (os.path.join(out_dir, f"{case['name']}.{e}") for e in ("item", "tr", "te"))
    flags = [x for x in case["flags"]]
    flags[flags.index("-t") + 1] = "20"          # two log lines (iterations 10 and 20)
    model = os.path.join(out_dir, "ref_model.txt")
    out = subprocess.check_output([os.path.join(HERE, "_ref", "train_ref")] + flags +
                                  ["-c", "1", "-p", te_p, "-o", model, item_p, tr_p], text=True)
    with open(os.path.join(out_dir, "ref_stdout.txt"), "w") as fh:
        fh.write(out)
    with open(os.path.join(out_dir, "cli_flags.txt"), "w") as fh:
        fh.write(" ".join(flags) + "\n")
    print(case["name"], "cli:\n" + out)


def run_ndcg_fixture():
    tool = os.path.join(REF, "script", "nDCG_degub_tool")
    out = subprocess.check_output(
        [os.path.join(HERE, "_ref", "train_ndcg_debug"), "-k", "8", "-t", "1", "-p",
         os.path.join(tool, "case1.mf"), os.path.join(tool, "test_item.mf"), os.path.join(tool, "case1.mf")],
        text=True)
    vals = []
    for line in out.splitlines():
        try:
            vals.append(float(line.strip()))
        except ValueError:
            pass
    labels = []
    with open(os.path.join(tool, "case1.mf")) as fh:
        for line in fh:
            labels.append([int(x) for x in line.split()[0].split(",")])
    n_items = sum(1 for _ in open(os.path.join(tool, "test_item.mf")))
    assert len(vals) == len(labels), (len(vals), len(labels))
    # the tool's own answer generator (gen_ans.py) restated as a cross-check of the binary's output
    import math
    for lab, v in zip(labels, vals):
        dcg = sum(1 / math.log2(i + 2) for i in range(10) if i in lab)
        idcg = sum(1 / math.log2(i + 2) for i in range(min(len(lab), 10)))
        assert abs(dcg / idcg - v) < 1e-4, (lab, v, dcg / idcg)
    with open(os.path.join(GOLD, "ndcg_case1.json"), "w") as fh:
        json.dump(dict(source="reference script/nDCG_degub_tool (case1.mf x test_item.mf), "
                              "train built with -DEBUG_nDCG -DSHOW_SCORE_ONLY: scores forced to z[i]=n-i, "
                              "prints per-user nDCG@10 with setprecision(4)",
                       n_items=n_items, ranking=list(range(n_items)), labels=labels, ndcg_at_10=vals), fh, indent=1)
    print("ndcg fixture:", vals)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    subprocess.check_call(["make", "-s", "-C", HERE, "ref"])
    for c in CASES:
        run_case(c)
    for c in CASES[:3]:
        run_cli(c)
    run_ndcg_fixture()
