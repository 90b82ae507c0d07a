"""CPU: the C-ABI library loads and exports every symbol include/vp8b200.h declares; with no
GPU it must refuse to create a context (there is no CPU fallback)."""
import ctypes
import os
import re

from conftest import ROOT
from vp8b200 import abi


def _declared():
    text = open(os.path.join(ROOT, "include", "vp8b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vp8b200_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = abi.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "libvp8b200.so does not export " + n
    assert sorted(abi.EXPORTS) == names
    assert L.vp8b200_abi_version() == 1


def test_no_cpu_fallback():
    L = abi.lib()
    if L.vp8b200_device_count() > 0:
        return                                   # on a GPU box the gpu tests cover creation
    h = ctypes.c_void_p()
    st = L.vp8b200_create(ctypes.byref(h), 0, 352, 288, 4)
    assert st == -2 and not h.value               # VP8B200_ERR_NO_DEVICE
    assert b"no CPU path" in L.vp8b200_strerror(st)


def test_argument_validation():
    L = abi.lib()
    h = ctypes.c_void_p()
    assert L.vp8b200_create(ctypes.byref(h), 0, 350, 288, 4) == -1      # not a multiple of 16
    assert L.vp8b200_create(ctypes.byref(h), 0, 352, 288, 99) == -1
    assert L.vp8b200_frame_submit(None, 0, 0) == -1
