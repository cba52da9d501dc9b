"""No-GPU checks of the drop-in boundary: libnnab.so loads and exports every
symbol include/nnab.h declares; the ctypes table covers exactly that set."""
import ctypes
import os
import re

from helpers import ROOT

from nnaudio_b200 import _C


def _declared():
    text = open(os.path.join(ROOT, "include", "nnab.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nnab_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 15
    handle = ctypes.CDLL(_C.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"libnnab.so does not export {n}"


def test_ctypes_table_matches_header():
    assert sorted(_C.SIGNATURES) == _declared()


def test_abi_version_and_strerror():
    lib = _C.lib()
    assert lib.nnab_abi_version() == 1
    assert lib.nnab_strerror(0) == b"ok"
    assert b"tcgen05" in lib.nnab_strerror(-2)
    assert lib.nnab_pack_tile_n(1025) == 208 and lib.nnab_pack_tile_n(84) == 176
