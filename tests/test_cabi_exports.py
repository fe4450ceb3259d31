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
    assert lib.vima_abi_version() == 5


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


def test_new_entry_points_refuse_cpu_inputs():
    """The rows added around the path (prepare, incremental decode, action post-processing, the baselines' encoders) have no
    CPU route either: CPU tensors raise instead of being computed by torch."""
    import numpy as np
    import torch

    import vima_b200
    from vima_b200 import nn as vnn
    from vima_b200.prepare import crop_objects

    if torch.cuda.is_available():
        pytest.skip("GPU box")
    with pytest.raises((RuntimeError, AssertionError)):
        crop_objects(np.zeros((1, 3, 8, 8), np.uint8), np.zeros((1, 8, 8), np.uint8), [1], device="cpu")
    pol = vima_b200.VIMAPolicy(embed_dim=256, xf_n_layers=1, sattn_n_heads=8, xattn_n_heads=8)
    with pytest.raises(RuntimeError, match="no CPU|CUDA"):
        pol.start_decode(torch.zeros(4, 1, 256), torch.ones(1, 4, dtype=torch.bool))
    with pytest.raises(RuntimeError, match="no CPU|CUDA"):
        pol.postprocess_actions({"pose0_position": torch.zeros(1, 1, 2, dtype=torch.int64)}, torch.zeros(1, 2), torch.ones(1, 2))
    enc = vnn.ObjectsPerceiverEncoder(64, num_latents=4, num_blocks=1, num_self_attends_per_block=1, num_self_attention_heads=8,
                                      num_cross_attention_heads=8, attention_probs_dropout_prob=0.1)
    with pytest.raises(RuntimeError, match="no CPU|CUDA"):
        enc(torch.zeros(2, 16, 64))


def _integration_md_class(name: str):
    """exec()s one `class <name>(C.Structure)` block out of INTEGRATION.md's reference-side stub."""
    import ctypes as C

    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"^class " + name + r"\(C\.Structure\):.*?\n(?=\S)", md, flags=re.S | re.M)
    assert m, f"class {name} not found in INTEGRATION.md"
    ns = {"C": C}
    exec(m.group(0), ns)
    return ns[name]


def test_integration_md_stub_matches_library():
    """The ctypes stub a reference maintainer would copy from INTEGRATION.md has the library's layout."""
    import ctypes as C

    import __graft_entry__

    __graft_entry__.build()
    from vima_b200 import _C

    lib = _C.load_library()
    doc = _integration_md_class("NormDesc")
    assert [f[0] for f in doc._fields_] == [f[0] for f in _C.NormDesc._fields_]
    assert [f[1] for f in doc._fields_] == [f[1] for f in _C.NormDesc._fields_]
    assert C.sizeof(doc) == lib.vima_sizeof_norm_desc() == C.sizeof(_C.NormDesc)
    assert doc._fields_[0][0] == "struct_size"


def test_ctypes_mirrors_match_header_offsets(tmp_path):
    """gcc compiles include/vima_b200.h as plain C and prints offsetof() of every descriptor field; the ctypes mirrors in
    vima_b200/_C.py (what the GPU tests call through) must agree field by field."""
    import ctypes as C
    import shutil
    import subprocess

    from vima_b200 import _C

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    pairs = [("vima_gemm_desc", _C.GemmDesc), ("vima_norm_desc", _C.NormDesc), ("vima_attn_desc", _C.AttnDesc), ("vima_f32_gemm_group", _C.F32GemmGroup)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "vima_b200.h")}"', "int main(void) {"]
    for cname, mirror in pairs:
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in mirror._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "offsets.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "offsets"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], check=True)
    got = dict(ln.split() for ln in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, mirror in pairs:
        assert int(got[cname]) == C.sizeof(mirror), cname
        for fname, _ in mirror._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(mirror, fname).offset, f"{cname}.{fname}"


@pytest.mark.gpu
def test_descriptor_struct_size_is_enforced():
    """A descriptor whose struct_size the library does not know is rejected before any field is read."""
    import ctypes as C

    import torch

    from vima_b200 import _C

    ctx = _C.Context.get(torch.device("cuda", 0))
    x = torch.randn(4, 64, device="cuda")
    out = torch.empty_like(x)
    for bad in (0, 8, C.sizeof(_C.NormDesc) + 8):
        d = _C.NormDesc()
        d.struct_size = bad
        d.x, d.rows, d.cols, d.ldx = x.data_ptr(), 4, 64, 64
        d.out_f32, d.ld_o32 = out.data_ptr(), 64
        rc = ctx.lib.vima_norm(ctx.h, C.byref(d), C.c_void_p(ctx._s()))
        assert rc == 1, bad  # VIMA_E_INVALID
        assert b"struct_size" in ctx.lib.vima_last_error(ctx.h)
    for desc, fn in ((_C.GemmDesc, ctx.lib.vima_gemm), (_C.AttnDesc, ctx.lib.vima_attention)):
        d = desc()
        d.struct_size = 4
        assert fn(ctx.h, C.byref(d), C.c_void_p(ctx._s())) == 1
