"""CPU: the C-ABI library loads and exports every symbol include/vima_b200.h declares (no compute without a GPU)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "vima_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vima_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    import __graft_entry__

    __graft_entry__.build()
    from vima_b200 import _C

    lib = _C.load_library()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vima_b200.h but not exported"
    assert sorted(_C.EXPORTS) == names
    assert lib.vima_abi_version() == 3


def test_no_cpu_fallback():
    """The product path refuses to run without an sm_100 device instead of silently falling back."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU box")
    from vima_b200 import _C

    with pytest.raises(RuntimeError):
        _C.Context.get(torch.device("cpu"))
    with pytest.raises(RuntimeError):
        _C.Context(0)
