"""CPU-side checks of the C-ABI boundary: the library loads and exports every symbol include/hyvae.h declares."""
import ctypes
import os
import re

import pytest

from hunyuanvideo_efficiency_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "hyvae.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hyvae_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_what_python_binds():
    assert set(_declared()) == set(N.EXPORTS)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(N.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(N.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name
    lib.hyvae_version.restype = ctypes.c_int
    assert lib.hyvae_version() == 121


def test_no_cpu_fallback():
    import torch
    from hunyuanvideo_efficiency_b200.vae import AutoencoderKLCausal3D
    from oracle import weights as W
    m = AutoencoderKLCausal3D.from_config(W.SMALL_CONFIG)
    with pytest.raises(N.HyvaeError):
        m.encode(torch.zeros(1, 3, 5, 32, 32))
