"""CPU tests: the C-ABI library loads without a GPU, exports every symbol the header declares,
its host-side planning (descriptor validation, workspace layout) behaves, and the Python
operator surface mirrors the reference's names / signatures / error behaviour."""
import ctypes
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pvnet_vote_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pvb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(pvb):
    lib = ctypes.CDLL(pvb._lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pvnet_vote_b200.h but not exported"
    assert set(names) == set(pvb._lib.SIGNATURES), "ctypes table and header disagree"
    assert pvb._lib.load().pvb_version() == 200


def _desc(pvb, **kw):
    d = pvb._lib.PvbDesc()
    base = dict(B=16, H=480, W=640, K=9, hn=512, inlier_thresh=0.99, min_num=5, max_num=30000,
                mask_dtype=pvb._lib.PVB_MASK_I64, select_mode=0)
    base.update(kw)
    for k, v in base.items():
        setattr(d, k, v)
    return d


def test_descriptor_layout_matches_c(pvb):
    # pvb_desc: 10 x 4 bytes, int64[3], int64[5], 2 x int32, uint64, 2 x int32
    assert ctypes.sizeof(pvb._lib.PvbDesc) == 128
    assert pvb._lib.PvbDesc.mask_stride.offset == 40
    assert pvb._lib.PvbDesc.seed.offset == 112


def test_workspace_layout(pvb):
    lib = pvb._lib.load()
    d = _desc(pvb)
    L = pvb._lib.PvbLayout()
    assert lib.pvb_workspace_layout(d, L) == 0
    assert L.nwords == 480 * 640 // 32
    assert L.capacity % 32 == 0 and 30000 < L.capacity < 32000      # max_num + 8 sigma + slack
    offs = [L.status, L.fgsum, L.nz, L.tn, L.state, L.ticket, L.blocktot, L.bits, L.xy, L.dirs, L.hyp, L.counts, L.win, L.total]
    assert offs == sorted(offs) and all(o % 256 == 0 for o in offs)
    assert L.dirs - L.xy >= 16 * L.capacity * 8
    assert L.hyp - L.dirs >= 16 * 9 * L.capacity * 8
    assert lib.pvb_workspace_bytes(d) == L.total
    # capacity never exceeds the image
    d2 = _desc(pvb, H=32, W=32)
    assert lib.pvb_workspace_layout(d2, L) == 0 and L.capacity == 1024
    d3 = _desc(pvb, capacity=100000)
    assert lib.pvb_workspace_layout(d3, L) == 0 and L.capacity == 100000


@pytest.mark.parametrize("bad", [dict(H=0), dict(K=0), dict(hn=0), dict(mask_dtype=9), dict(select_mode=2),
                                 dict(B=70000), dict(H=40000, W=40000)])
def test_invalid_descriptors_are_rejected(pvb, bad):
    lib = pvb._lib.load()
    L = pvb._lib.PvbLayout()
    assert lib.pvb_workspace_layout(_desc(pvb, **bad), L) == pvb._lib.PVB_ERR_INVALID
    assert lib.pvb_workspace_bytes(_desc(pvb, **bad)) == 0
    assert len(lib.pvb_last_error()) > 0


def test_null_and_small_workspace_errors(pvb):
    lib = pvb._lib.load()
    d = _desc(pvb, B=1, H=8, W=8, K=1, hn=4)
    # argument validation happens before any CUDA call, so it is testable without a GPU
    assert lib.pvb_ransac_voting_v3(d, None, None, None, None, None, None, 0, None) == pvb._lib.PVB_ERR_INVALID
    buf = ctypes.create_string_buffer(1024)
    rc = lib.pvb_ransac_voting_v3(d, ctypes.addressof(buf), ctypes.addressof(buf), None, None, ctypes.addressof(buf),
                                  ctypes.addressof(buf), 16, None)
    assert rc == pvb._lib.PVB_ERR_WORKSPACE
    with pytest.raises(RuntimeError, match="workspace"):
        pvb._lib.check(rc)


def test_operator_surface_matches_reference_signatures(pvb):
    """Names, positional order and defaults of ransac_voting_gpu.py:6-7, :112-113, :202."""
    g = pvb.ransac_voting_gpu

    def positional(fn):
        return [(p.name, p.default) for p in inspect.signature(fn).parameters.values()
                if p.kind == p.POSITIONAL_OR_KEYWORD]

    ref_v3 = [("mask", inspect._empty), ("vertex", inspect._empty), ("round_hyp_num", inspect._empty),
              ("inlier_thresh", 0.999), ("confidence", 0.99), ("max_iter", 20), ("min_num", 5), ("max_num", 30000)]
    assert positional(g.ransac_voting_layer_v3) == ref_v3
    assert positional(g.ransac_voting_layer) == ref_v3
    assert positional(g.estimate_voting_distribution_with_mean) == [
        ("mask", inspect._empty), ("vertex", inspect._empty), ("mean", inspect._empty), ("round_hyp_num", 256),
        ("min_hyp_num", 4096), ("topk", 128), ("inlier_thresh", 0.99), ("min_num", 5), ("max_num", 30000),
        ("output_hyp", False)]
    for n in ("generate_hypothesis", "voting_for_hypothesis", "generate_hypothesis_vanishing_point",
              "voting_for_hypothesis_vanishing_point"):
        assert callable(getattr(pvb.ransac_voting, n))


def test_cpu_tensors_are_rejected_like_the_reference(pvb):
    """ransac_voting.cpp:7-9 asserts CUDA + contiguous; there is no CPU path here either."""
    mask = torch.ones(1, 8, 8, dtype=torch.int64)
    vertex = torch.zeros(1, 8, 8, 2, 2)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        pvb.ransac_voting_layer_v3(mask, vertex, 8)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        pvb.estimate_voting_distribution_with_mean(mask, vertex, torch.zeros(1, 2, 2))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        pvb.ransac_voting.generate_hypothesis(torch.zeros(4, 1, 2), torch.zeros(4, 2), torch.zeros(2, 1, 2, dtype=torch.int32))


def test_install_as_reference_module(pvb):
    import sys
    pvb.install_as_reference_module()
    from lib.csrc.ransac_voting.ransac_voting_gpu import (ransac_voting_layer, ransac_voting_layer_v3,  # noqa: F401
                                                          estimate_voting_distribution_with_mean)
    import lib.csrc.ransac_voting.ransac_voting as ext
    assert ext is pvb.ransac_voting
    assert ransac_voting_layer_v3 is pvb.ransac_voting_layer_v3
    for k in [k for k in sys.modules if k == "lib" or k.startswith("lib.")]:
        del sys.modules[k]


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "clean_pvnet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pvnet_oracle" not in text and "oracle/" not in text.replace("see oracle/", ""), f


def test_missing_library_fails_loudly(pvb, monkeypatch):
    """No CPU / PyTorch fallback: without the CUDA library the product raises, it never degrades."""
    monkeypatch.setattr(pvb._lib, "_LIB", None)
    monkeypatch.setattr(pvb._lib, "LIB_PATH", "/nonexistent/libpvnet_vote_b200.so")
    with pytest.raises(RuntimeError, match="no CPU/PyTorch fallback"):
        pvb._lib.load()


def test_plain_c_consumer_links_and_runs(pvb, tmp_path):
    """include/pvnet_vote_b200.h is C99 and the shared library needs nothing but libc/libstdc++ at link time."""
    import subprocess
    exe = str(tmp_path / "cabi_smoke")
    libdir = os.path.dirname(pvb._lib.LIB_PATH)
    cc = os.environ.get("CC", "gcc")
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cabi_smoke.c"), "-L", libdir, "-lpvnet_vote_b200",
                           "-Wl,-rpath," + libdir, "-o", exe])
    r = subprocess.run([exe], stdout=subprocess.PIPE, text=True, timeout=60)
    assert r.returncode == 0, r.stdout
    assert "sizeof(pvb_desc)=128" in r.stdout


def test_exchange_entry_points_fail_cleanly_without_a_gpu(pvb):
    """pvb_exchange_*: argument validation runs before any CUDA call, and without a usable device creation returns
    PVB_ERR_CUDA with a message instead of crashing (there is no CPU implementation of the exchange either)."""
    lib = pvb._lib.load()
    h = ctypes.c_void_p()
    assert lib.pvb_exchange_create(3, 2, 4, 64, ctypes.byref(h)) == pvb._lib.PVB_ERR_INVALID and not h.value
    assert lib.pvb_exchange_create(0, 17, 4, 64, ctypes.byref(h)) == pvb._lib.PVB_ERR_INVALID        # > PVB_MAX_PEERS
    assert lib.pvb_exchange_create(0, 2, 1, 64, ctypes.byref(h)) == pvb._lib.PVB_ERR_INVALID         # slots < 2
    assert lib.pvb_exchange_create(0, 2, 4, 0, ctypes.byref(h)) == pvb._lib.PVB_ERR_INVALID
    assert b"bytes_per_rank" in lib.pvb_last_error()
    if not torch.cuda.is_available():
        rc = lib.pvb_exchange_create(0, 2, 4, 64, ctypes.byref(h))
        assert rc == pvb._lib.PVB_ERR_CUDA and not h.value and lib.pvb_last_error()
    assert lib.pvb_exchange_wait(None, 1, None, None, 1.0, None) == pvb._lib.PVB_ERR_INVALID
    assert lib.pvb_exchange_destroy(None) == pvb._lib.PVB_OK
