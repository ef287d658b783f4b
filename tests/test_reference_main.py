"""The drop-in claim of INTEGRATION.md section A, checked: the reference's OWN driver
(/root/reference/train.cpp, never copied into this repo) compiles unchanged against the mirror
header one-class-ffm_b200/host/ffm.h and links with the B200 host layer + libocffm_cuda.so.

CPU part (this container, where /root/reference exists): compile + link from the sources where
they lie, the usage text, and the loud failure without a device.  GPU part: the binary built by
`make -C one-class-ffm_b200/host` (train_refmain, it travels to the GPU box as a built artefact)
reproduces the reference's log byte for byte and its model file on tests/golden/tiny*."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from test_train_cli import read_model

HOST = os.path.join(ROOT, "one-class-ffm_b200", "host")
PKG = os.path.join(ROOT, "one-class-ffm_b200")
REF_MAIN = "/root/reference/train.cpp"
BIN = os.path.join(PKG, "train_refmain")


@pytest.mark.skipif(not os.path.exists(REF_MAIN), reason="needs /root/reference (builder container only)")
def test_reference_driver_compiles_and_links_unchanged():
    with tempfile.TemporaryDirectory() as tmp:
        exe, obj = os.path.join(tmp, "train_refmain"), os.path.join(tmp, "refmain.o")
        # from stdin: `#include "ffm.h"` must resolve to the mirror, not to the reference's own header
        # next to train.cpp (quote includes search the including file's directory first)
        with open(REF_MAIN) as src:
            r = subprocess.run(["g++", "-O1", "-std=c++17", "-fopenmp", "-iquote", HOST, "-x", "c++", "-c", "-o", obj, "-"],
                               stdin=src, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-4000:]
        r = subprocess.run(["g++", "-O1", "-std=c++17", "-fopenmp", "-I", HOST, "-o", exe, obj,
                            os.path.join(HOST, "ffm_host.cpp"), "-L", PKG, "-locffm_cuda", f"-Wl,-rpath,{PKG}"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-4000:]
        # the reference's usage text comes from its own main (train.cpp:34-51)
        u = subprocess.run([exe], capture_output=True, text=True)
        assert u.returncode == 1 and "usage: train [options] item_feature_file train_file" in u.stdout + u.stderr
        import ocffm
        if ocffm.device_count() == 0:
            # no CPU fallback behind the reference's entry points either
            base = os.path.join(GOLDEN, "tiny", "tiny")
            f = subprocess.run([exe, "-k", "8", "-t", "1", base + ".item", base + ".tr"],
                               capture_output=True, text=True)
            assert f.returncode != 0
            assert "no CUDA device" in f.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["tiny", "tiny_ns", "tiny_freq"])
def test_reference_driver_on_b200_matches_reference_log(case):
    if not os.path.exists(BIN):
        pytest.skip("train_refmain not built (needs /root/reference at build time)")
    gdir = os.path.join(GOLDEN, case)
    flags = open(os.path.join(gdir, "cli_flags.txt")).read().split()
    want_out = open(os.path.join(gdir, "ref_stdout.txt")).read()
    base = os.path.join(gdir, case)
    with tempfile.TemporaryDirectory() as tmp:
        model = os.path.join(tmp, "model.txt")
        r = subprocess.run([BIN] + flags + ["-c", "1", "-p", base + ".te", "-o", model, base + ".item", base + ".tr"],
                           capture_output=True, text=True, env=dict(os.environ, OCFFM_DTYPE="f64"))
        assert r.returncode == 0, r.stderr
        assert r.stdout == want_out                      # byte-identical log, reference main + B200 solver
        got_hdr, got_rows = read_model(model)
    want_hdr, want_rows = read_model(os.path.join(gdir, "ref_model.txt"))
    assert got_hdr == want_hdr and list(got_rows) == list(want_rows)
    num = max(np.max(np.abs(got_rows[k] - want_rows[k])) for k in want_rows)
    den = max(np.max(np.abs(v)) for v in want_rows.values())
    assert num <= 2e-5 * den
