"""End-to-end drop-in check: the B200 `train` binary against the unmodified reference CLI on the
same files and flags.  Both draw the same initial model (same rand()/libstdc++ stream), so the
fp64 device path must print the reference's log lines and write its model file."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
TRAIN = os.path.join(ROOT, "one-class-ffm_b200", "train")
NUM = re.compile(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?")


def numbers(text):
    return np.array([float(x) for x in NUM.findall(text)])


def read_model(path):
    rows = {}
    with open(path) as fh:
        lines = fh.read().split("\n")
    hdr = []
    for ln in lines:
        if not ln:
            continue
        if ln[0] in "WH":
            key, *vals = ln.split(" ")
            rows[key] = np.array([float(v) for v in vals])
        else:
            hdr.append(int(ln))
    return hdr, rows


@pytest.mark.parametrize("case", ["tiny", "tiny_ns", "tiny_freq"])
@pytest.mark.parametrize("mode", ["f64", "f32"])
def test_train_matches_reference_cli(case, mode):
    gdir = os.path.join(GOLDEN, case)
    flags = open(os.path.join(gdir, "cli_flags.txt")).read().split()
    want_out = open(os.path.join(gdir, "ref_stdout.txt")).read()
    base = os.path.join(gdir, case)
    with tempfile.TemporaryDirectory() as tmp:
        model = os.path.join(tmp, "model.txt")
        extra = ["--f64"] if mode == "f64" else []
        r = subprocess.run([TRAIN] + flags + extra + ["-c", "1", "-p", base + ".te", "-o", model,
                                                       base + ".item", base + ".tr"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        got_hdr, got_rows = read_model(model)
    want_hdr, want_rows = read_model(os.path.join(gdir, "ref_model.txt"))
    # identical layout: header line, the iteration numbers, same count of printed fields
    assert r.stdout.split("\n")[0] == want_out.split("\n")[0]
    g, w = numbers(r.stdout), numbers(want_out)
    assert g.shape == w.shape
    if mode == "f64":
        assert r.stdout == want_out                      # byte-identical log
        tol = 2e-5                                       # model text has 6 significant digits
    else:
        assert np.allclose(g, w, rtol=2e-2, atol=0.06)   # 3 printed digits, fp32 solver drift
        tol = 5e-2
    assert got_hdr == want_hdr
    assert list(got_rows) == list(want_rows)             # same rows in the same order
    num = max(np.max(np.abs(got_rows[k] - want_rows[k])) for k in want_rows)
    den = max(np.max(np.abs(v)) for v in want_rows.values())
    assert num <= tol * den, (num, den)


def test_grid_runner_matches_separate_runs():
    """--grid-l / --grid-w solve every (lambda, omega) point on the data uploaded once; each point
    must print exactly the log of a separate run with those flags (fp64: byte for byte) and write
    the same model."""
    gdir = os.path.join(GOLDEN, "tiny")
    flags = open(os.path.join(gdir, "cli_flags.txt")).read().split()
    base = os.path.join(gdir, "tiny")

    def strip(flag, argv):   # drop "-l v" / "-w v" from the golden flags
        out, skip = [], False
        for a in argv:
            if skip:
                skip = False
            elif a == flag:
                skip = True
            else:
                out.append(a)
        return out
    common = strip("-w", strip("-l", flags)) + ["--f64", "-c", "1", "-p", base + ".te"]
    data = [base + ".item", base + ".tr"]
    ls, ws = ["4", "0.5"], ["0.0078125", "0.0625"]
    with tempfile.TemporaryDirectory() as tmp:
        gm = os.path.join(tmp, "grid_model.txt")
        g = subprocess.run([TRAIN] + common + ["--grid-l", ",".join(ls), "--grid-w", ",".join(ws), "-o", gm] + data,
                           capture_output=True, text=True)
        assert g.returncode == 0, g.stderr
        sections = g.stdout.split("config ")[1:]
        assert len(sections) == 4
        i = 0
        for l in ls:
            for w in ws:
                head, _, body = sections[i].partition("\n")
                assert head == f"-l {l} -w {w}"
                sm = os.path.join(tmp, f"single_{i}.txt")
                one = subprocess.run([TRAIN] + common + ["-l", l, "-w", w, "-o", sm] + data,
                                     capture_output=True, text=True)
                assert one.returncode == 0, one.stderr
                assert body == one.stdout, (l, w)
                _, got = read_model(f"{gm}.l{l}.w{w}")
                _, want = read_model(sm)
                assert list(got) == list(want)
                # 6 printed digits; fp64 atomics may reorder sums, so allow the last digit to move
                assert all(np.allclose(got[k], want[k], rtol=1e-5, atol=1e-9) for k in want), (l, w)
                i += 1


@pytest.mark.parametrize("case", ["tiny", "tiny_ns"])
def test_binary_model_save_load_predict_round_trip(case):
    """`--save-binary` writes the reference's binary layout (ffm.cpp:1239-1267);
    `--load <file> --predict-only` must then print the metrics the training run printed for its
    last iteration (same model, same validate), and `--load` + more passes must continue from it."""
    gdir = os.path.join(GOLDEN, case)
    flags = open(os.path.join(gdir, "cli_flags.txt")).read().split()
    base = os.path.join(gdir, case)
    data = ["-c", "1", "-p", base + ".te", base + ".item", base + ".tr"]
    with tempfile.TemporaryDirectory() as tmp:
        mb, mt = os.path.join(tmp, "m.bin"), os.path.join(tmp, "m.txt")
        a = subprocess.run([TRAIN] + flags + ["--f64", "--save-binary", mb, "-o", mt] + data, capture_output=True, text=True)
        assert a.returncode == 0, a.stderr
        last = [ln for ln in a.stdout.split("\n") if ln.strip()][-1]
        mt2 = os.path.join(tmp, "m2.txt")
        b = subprocess.run([TRAIN] + flags + ["--f64", "--load", mb, "--predict-only", "-o", mt2] + data,
                           capture_output=True, text=True)
        assert b.returncode == 0, b.stderr
        pred = [ln for ln in b.stdout.split("\n") if ln.strip()][-1]
        # same columns after the iteration number
        assert numbers(last)[1:].tolist() == numbers(pred)[1:].tolist(), (last, pred)
        assert open(mt).read() == open(mt2).read()       # the text model of the reloaded state is the same file
        # the binary file itself: u32 f, fu, fv, k then u64 Ds (reference layout)
        hdr = np.fromfile(mb, dtype=np.uint32, count=4)
        want_hdr, _ = read_model(mt)
        assert list(hdr) == want_hdr[:4]


def test_gpu_init_flag_trains_and_is_reproducible():
    """--gpu-init <seed>: the initial model is drawn on the device (ocffm_init_model); two runs with the
    same seed print the same log and write the same model, another seed gives another model."""
    gdir = os.path.join(GOLDEN, "tiny")
    flags = open(os.path.join(gdir, "cli_flags.txt")).read().split()
    base = os.path.join(gdir, "tiny")
    outs = []
    with tempfile.TemporaryDirectory() as tmp:
        for i, seed in enumerate(["11", "11", "12"]):
            model = os.path.join(tmp, f"m{i}.txt")
            r = subprocess.run([TRAIN] + flags + ["--f64", "--gpu-init", seed, "-c", "1", "-p", base + ".te", "-o", model,
                                                   base + ".item", base + ".tr"], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
            outs.append((r.stdout, open(model).read()))
    assert outs[0] == outs[1]
    assert outs[0][1] != outs[2][1]
    assert len(outs[0][0].strip().split("\n")) == 3          # header + iterations 10 and 20
