"""The C-ABI library loads and exports every symbol include/kwb200.h declares; host-only entry points behave.
No compute calls (no GPU here)."""
import ctypes
import os
import re

import numpy as np

from kotoba_whisper_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "kwb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kw_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in kwb200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)
    assert b"sm_100a" in lib.kw_version()


def test_struct_sizes_match_header_layout():
    assert ctypes.sizeof(_lib.kw_config) == 11 * 4
    assert ctypes.sizeof(_lib.kw_enc_layer_weights) == 12 * 8
    assert ctypes.sizeof(_lib.kw_dec_layer_weights) == 20 * 8
    assert ctypes.sizeof(_lib.kw_weights) == 13 * 8
    assert ctypes.sizeof(_lib.kw_token_rules) == 4 * 4 + 8 + 8 + 8 + 8  # pointers 8-aligned, trailing pad


def test_host_filterbank_matches_oracle():
    from oracle.logmel_ref import mel_filter_bank
    lib = _lib.load()
    for nm in (80, 128):
        out = np.zeros((201, nm))
        assert lib.kw_mel_filterbank(nm, out.ctypes.data) == 0
        ref = mel_filter_bank(nm)
        assert np.abs(out - ref).max() < 1e-15
        assert ((out != 0) == (ref != 0)).all()


def test_argument_errors_are_reported_not_raised():
    lib = _lib.load()
    assert lib.kw_mel_filterbank(0, None) == -1
    assert b"kw_mel_filterbank" in lib.kw_last_error()
    assert lib.kw_logmel(None, None, 1, 480000, 128, None, None, None) == -1
    assert lib.kw_encode(None, None, 1, None, None) == -1
    assert lib.kw_launch_count(1) >= 0


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "kotoba_whisper_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
