"""CPU: the C-ABI library builds, loads and exports every symbol include/said_b200.h declares.
No compute calls (there is no GPU here); on a GPU-less box the library must fail loudly, not fall back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "said_b200.h")).read()
    return sorted(set(re.findall(r"SAID_API[^;(]*?\b(said_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    from said_b200 import _lib

    assert declared_symbols() == sorted(_lib.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    from said_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    lib = _lib.load_library()
    for s in declared_symbols():
        assert hasattr(lib, s), s
    assert lib.said_version() == 1


def test_denoise_args_struct_layout():
    """ctypes mirror of said_denoise_args: field order/types as declared in the header."""
    from said_b200 import _lib

    text = open(os.path.join(ROOT, "include", "said_b200.h")).read()
    body = re.search(r"typedef struct said_denoise_args \{(.*?)\} said_denoise_args;", text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            names.append(re.findall(r"(\w+)\s*$", part.strip())[0])
    assert names == [f[0] for f in _lib.DenoiseArgs._fields_]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly():
    from said_b200 import _lib

    lib = _lib.load_library()
    h = ctypes.c_void_p()
    rc = lib.said_create(0, ctypes.byref(h))
    assert rc != 0 and h.value is None
    assert len(lib.said_last_error()) > 0
    with pytest.raises(_lib.SaidLibraryError):
        _lib.Engine(torch.device("cpu"))
