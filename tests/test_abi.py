"""The C-ABI library loads, exports exactly what include/b200rnn.h declares, and its host-side logic
(descriptor validation, workspace sizing, error reporting) works without a GPU."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from b200rnn import _lib


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200rnn.h")).read()
    return sorted(set(re.findall(r"B200RNN_API\s+[\w\s\*]+?\b(b200rnn_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for sym in _declared_symbols():
        assert hasattr(lib, sym), sym
    assert _lib.load().b200rnn_version() == _lib.ABI_VERSION


def test_workspace_bytes_scale_with_the_problem():
    small = _lib.workspace_bytes(_lib.Desc(_lib.GRU, 8, 3, 256, 256, 2, 1, 1, 0.5, 0))
    big = _lib.workspace_bytes(_lib.Desc(_lib.GRU, 128, 120, 256, 256, 2, 1, 1, 0.5, 0))
    assert all(b > s > 0 for s, b in zip(small, big))
    # reserve holds, per layer: gates [T,B,3H] + hn [T,B,H]; plus layer-0 output raw and dropped [T,B,H]
    T, B, H = 120, 128, 256
    expect = 4 * T * B * (2 * 4 * H + 2 * H)
    assert expect <= big[0] <= expect * 1.01 + 4096
    lstm = _lib.workspace_bytes(_lib.Desc(_lib.LSTM, 64, 30, 1024, 256, 2, 2, 1, 0.0, 0))
    assert lstm[0] >= 4 * 30 * 64 * (2 * 2 * 5 * 256 + 2 * 256)


@pytest.mark.parametrize("desc,frag", [
    (_lib.Desc(7, 4, 4, 16, 128, 1, 1, 0, 0.0, 0), "mode"),
    (_lib.Desc(_lib.GRU, 4, 4, 16, 100, 1, 1, 0, 0.0, 0), "hidden_size"),
    (_lib.Desc(_lib.GRU, 4, 4, 16, 128, 1, 3, 0, 0.0, 0), "bad shape"),
    (_lib.Desc(_lib.LSTM, 4, 4, 0, 128, 1, 1, 0, 0.0, 0), "bad shape"),
    (_lib.Desc(_lib.LSTM, 4, 4, 16, 128, 1, 1, 0, 1.5, 0), "dropout_p"),
])
def test_invalid_descriptors_are_rejected_with_a_message(desc, frag):
    with pytest.raises(_lib.B200RNNError) as ei:
        _lib.workspace_bytes(desc)
    assert frag in str(ei.value)


def test_forward_rejects_null_pointers_before_touching_the_device():
    lib = _lib.load()
    d = _lib.Desc(_lib.GRU, 2, 2, 16, 128, 1, 1, 0, 0.0, 0)
    rc = lib.b200rnn_forward(ctypes.byref(d), None, 0, 0, None, None, 0, 0, None, None, None, None, 0, 0, None, None)
    assert rc == -1 and b"null pointer" in lib.b200rnn_last_error()
