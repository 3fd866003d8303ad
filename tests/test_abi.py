"""The C-ABI library loads and exports every symbol include/mis_b200.h declares (no GPU compute)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mis_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mis_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    from medical_image_segmentation_b200 import _lib
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in mis_b200.h but not exported by libmis_b200.so"
        assert n in _lib.EXPORTS, f"{n} has no ctypes signature in _lib.EXPORTS"
    assert sorted(_lib.EXPORTS) == names


def test_version_and_struct_layout():
    from medical_image_segmentation_b200 import _lib
    assert _lib.lib.mis_version() == 1
    src = open(HEADER).read()
    assert "#define MIS_ABI_VERSION 1" in src
    d = _lib.VIEW_PARAMS_DTYPE
    assert d.itemsize == 48
    assert [d.fields[k][1] for k in ("img", "top", "left", "h", "w", "flags", "order", "brightness", "contrast",
                                     "saturation", "hue", "blur_sigma")] == [0, 4, 8, 12, 16, 20, 24, 28, 32, 36, 40, 44]


def test_argument_errors_without_gpu():
    """Validation happens before any CUDA call, so these run on a CPU-only box."""
    from medical_image_segmentation_b200 import _lib
    lib = _lib.lib
    assert lib.mis_aug_two_view(None, 1, 1, 8, 8, 64, None, 1, 0.0, 65535.0, None, None, None, 8, 0, 1, None) == _lib.MIS_ERR_INVALID_ARG
    assert b"null" in lib.mis_last_error()
    buf = np.zeros(64, np.uint8)
    f = (C.c_float * 1)(0.5)
    p = buf.ctypes.data
    args = lambda **kw: [kw.get("src", p), 1, kw.get("C", 1), 8, kw.get("W", 8), 64, p, 1, 0.0, kw.get("hi", 65535.0),
                         C.cast(f, C.c_void_p), C.cast(f, C.c_void_p), p, kw.get("s", 8), kw.get("dt", 0), 1, None]
    assert lib.mis_aug_two_view(*args(C=2)) == _lib.MIS_ERR_UNSUPPORTED
    assert lib.mis_aug_two_view(*args(C=3, s=12)) == _lib.MIS_ERR_UNSUPPORTED      # 3 channels: crop a multiple of 8
    assert lib.mis_aug_two_view(*args(W=7)) == _lib.MIS_ERR_UNSUPPORTED
    assert lib.mis_aug_two_view(*args(s=300)) == _lib.MIS_ERR_UNSUPPORTED
    assert lib.mis_aug_two_view(*args(dt=7)) == _lib.MIS_ERR_INVALID_ARG
    assert lib.mis_aug_two_view(*args(hi=0.0)) == _lib.MIS_ERR_INVALID_ARG
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.MIS_ERR_UNSUPPORTED, "x")
    with pytest.raises(ValueError):
        _lib.check(_lib.MIS_ERR_INVALID_ARG, "x")
    assert lib.mis_ntxent_fwd(p, 100, 64, 0, 100, 10.0, p, p, p, 1 << 30, None) == _lib.MIS_ERR_INVALID_ARG   # cols not padded
    assert lib.mis_ntxent_fwd(p, 128, 64, 0, 99, 10.0, p, p, p, 1 << 30, None) == _lib.MIS_ERR_INVALID_ARG    # odd rows
    assert lib.mis_ntxent_padded_rows(100) == 128 and lib.mis_ntxent_padded_rows(256) == 256
    assert lib.mis_ntxent_fwd(p, 128, 64, 0, 128, 100.0, p, p, p, 1 << 30, None) == _lib.MIS_ERR_UNSUPPORTED
    assert lib.mis_ntxent_fwd(p, 128, 64, 0, 128, 10.0, p, p, p, 16, None) == _lib.MIS_ERR_INVALID_ARG
    assert lib.mis_ntxent_scratch_bytes(2048, 2048, 128) > 2048 * 128 * 4
    assert lib.mis_draw_two_view_params(p, 10, 1, 0, 8, 8, C.cast(f, C.c_void_p), C.cast(f, C.c_void_p), p, p) == _lib.MIS_ERR_INVALID_ARG


def test_algorithmic_bytes_matches_formula():
    import torch
    from medical_image_segmentation_b200 import FusedTwoViewTransforms, algorithmic_bytes
    t = FusedTwoViewTransforms(224, (0.2,), (0.2,))
    torch.manual_seed(0)
    p = t.draw_params(64, 512, 512)
    want = int((2 * p["h"].astype(np.int64) * p["w"]).sum() + 2 * 224 * 224 * len(p))
    assert algorithmic_bytes(p, 1, 224) == want
    assert algorithmic_bytes(p, 1, 224, torch.float32) == want + 2 * 224 * 224 * len(p)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "medical_image_segmentation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_peer_exchange_argument_errors_and_mode_switch(monkeypatch):
    """The NVLink-exchange entry points validate their peer tables before any CUDA call; MIS_NTXENT_EXCHANGE=nccl
    turns the peer path off without touching symmetric memory."""
    from medical_image_segmentation_b200 import _lib, peer
    lib = _lib.lib
    buf = np.zeros(256, np.uint8)
    p = buf.ctypes.data
    tbl = (C.c_void_p * 8)(*([p] * 8))
    null_tbl = (C.c_void_p * 8)(*([p] + [None] * 7))

    def fwd(world, rank, u0=tbl, timeout=600.0, rows=128):
        return lib.mis_ntxent_fwd_peer(p, 1, rows, 64, 10.0, world, rank, u0, tbl, tbl, tbl, tbl, timeout, p, p, p, 1 << 30, None)

    assert fwd(9, 0) == _lib.MIS_ERR_INVALID_ARG            # > 8 ranks
    assert fwd(2, 2) == _lib.MIS_ERR_INVALID_ARG            # rank >= world
    assert fwd(1, 0) == _lib.MIS_ERR_INVALID_ARG            # a single rank has nothing to exchange (mis_ntxent_fwd_bwd)
    assert fwd(2, 0, u0=null_tbl) == _lib.MIS_ERR_INVALID_ARG
    assert b"rank 1" in lib.mis_last_error()
    assert fwd(2, 0, timeout=0.0) == _lib.MIS_ERR_INVALID_ARG
    assert fwd(2, 0, rows=99) == _lib.MIS_ERR_INVALID_ARG    # rows = [view 1; view 2] must be even
    assert lib.mis_ntxent_bwd_peer(p, 1, p, 128, 64, 10.0, 1.0, None, p, 2, 5, tbl, tbl, tbl, tbl, tbl, 600.0, p, 1 << 30,
                                   None) == _lib.MIS_ERR_INVALID_ARG
    monkeypatch.setenv("MIS_NTXENT_EXCHANGE", "nccl")
    assert peer.mode() == "nccl" and peer.get_exchange(None, 128, 64, None) is None
    monkeypatch.delenv("MIS_NTXENT_EXCHANGE")
    assert peer.mode() == "auto"
    assert peer._CTL_BYTES >= 4 * (2 * peer.MAX_PEERS + 4)        # PeerCtl: flags [2][8], counter, epoch, abort, pad
    assert peer.timeout_s() == 600.0
    monkeypatch.setenv("MIS_PEER_TIMEOUT_S", "30")
    assert peer.timeout_s() == 30.0


def test_h2d_staging_argument_errors_without_gpu():
    """mis_h2d_needed_rows validates sizes and every record before the first copy is enqueued."""
    from medical_image_segmentation_b200 import _lib
    lib = _lib.lib
    buf = np.zeros(2 * 8 * 8, np.uint16)
    rec = np.zeros(2, _lib.VIEW_PARAMS_DTYPE)
    rec["img"] = (0, 1)
    rec["h"] = rec["w"] = 4
    p, r = buf.ctypes.data, rec.ctypes.data
    assert lib.mis_h2d_needed_rows(None, p, 2, 1, 8, 8, 64, r, 2, 0, None, None) == _lib.MIS_ERR_INVALID_ARG
    assert lib.mis_h2d_needed_rows(p, p, 2, 1, 8, 8, 32, r, 2, 0, None, None) == _lib.MIS_ERR_INVALID_ARG        # stride < C*H*W
    rec["top"] = (0, 6)                                                                                          # rows [6, 10) of 8
    assert lib.mis_h2d_needed_rows(p, p, 2, 1, 8, 8, 64, r, 2, 0, None, None) == _lib.MIS_ERR_INVALID_ARG
    assert b"record 1" in lib.mis_last_error()
    rec["top"] = 0
    rec["img"] = (0, 2)                                                                                          # slice 2 of 2
    assert lib.mis_h2d_needed_rows(p, p, 2, 1, 8, 8, 64, r, 2, 0, None, None) == _lib.MIS_ERR_INVALID_ARG
