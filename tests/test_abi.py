"""The C-ABI library loads and exports every symbol include/*.h declares (no compute calls)."""
import ctypes
import os
import re

from wolkenbase_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(wb_[a-z0-9_]+)\s*\(", txt)))


def test_wolken_exports():
    names = _declared("wolken_b200.h")
    assert len(names) >= 30
    L = ctypes.CDLL(api.LIB_PATH)
    for n in names:
        assert hasattr(L, n), n
    assert sorted(api.EXPORTS) == names


def test_synth_exports():
    L = synth.lib()
    for n in _declared("wb_synth.h"):
        assert hasattr(L, n), n


def test_host_helpers_need_no_gpu():
    # sizeFit / bbox cube / ldecimal are host arithmetic and must agree with the oracle
    import numpy as np
    from oracle import wb_oracle as O
    L = api.lib()
    corners = np.array([[500000.0, 4200000.0, 120.249], [500054.655, 4200054.655, 133.94]])
    c1, s1 = (ctypes.c_double * 3)(), ctypes.c_double()
    c2, s2 = (ctypes.c_double * 3)(), ctypes.c_double()
    L.wb_size_fit(corners.ctypes.data, 2, c1, ctypes.byref(s1))
    O.lib().wbo_size_fit(corners.ctypes.data, 2, c2, ctypes.byref(s2))
    assert list(c1) == list(c2) and s1.value == s2.value
    q1, q2 = (ctypes.c_double * 4)(), (ctypes.c_double * 4)()
    L.wb_bbox_cube(corners.ctypes.data, 2, q1)
    O.lib().wbo_bbox_cube(corners.ctypes.data, 2, q2)
    assert list(q1) == list(q2)
    for x in [5e5, 42e5, 0.25, 124.0, 1e-7, -3.5, 123456.789]:
        assert api.ldecimal(x) == O.ldecimal(x)


def test_no_gpu_means_loud_failure():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.WolkenError):
        api.Context(0)
