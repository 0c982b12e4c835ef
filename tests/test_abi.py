"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/*.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "slfp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(slfp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from cnns_slfp_quantization_b200 import build, _native
    path = build.build()
    handle = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(handle, n), n
    assert sorted(_native.EXPORTED_SYMBOLS) == names
    assert _native.lib().slfp_version() == 200


def test_struct_layout_matches_header():
    from cnns_slfp_quantization_b200 import _native
    assert ctypes.sizeof(_native.SlfpConvDesc) == 19 * 4
    # pointers 8-byte aligned, ints/floats 4: computed by the C compiler for the same field order
    import subprocess, tempfile, textwrap
    src = textwrap.dedent('''
        #include <stdio.h>
        #include <stddef.h>
        #include "slfp_b200.h"
        int main(void) { printf("%zu %zu %zu %zu %zu", sizeof(SlfpEpilogue), offsetof(SlfpEpilogue, y_f32),
                                offsetof(SlfpEpilogue, next_k_div), offsetof(SlfpEpilogue, y_codes2),
                                offsetof(SlfpEpilogue, residual_f16)); return 0; }''')
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        got = [int(v) for v in subprocess.check_output([exe]).split()]
    E = _native.SlfpEpilogue
    assert got == [ctypes.sizeof(E), E.y_f32.offset, E.next_k_div.offset, E.y_codes2.offset, E.residual_f16.offset]


def test_product_never_imports_the_oracle():
    """No file of the product package imports, links or executes anything under oracle/."""
    pkg = os.path.join(ROOT, "cnns_slfp_quantization_b200")
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle|#include\s+[<\"].*oracle)|oracle[/\\.](slfp_oracle|torch_port|_build|_ref)", re.M)
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dp, f), errors="ignore").read()
                hits = [m.group(0) for m in pat.finditer(text)]
                # docstrings may NAME oracle/torch_port.py as what tests pass in; code may not touch it
                code_hits = [h for h in hits if h.lstrip().startswith(("from", "import", "#include"))]
                assert not code_hits, (os.path.join(dp, f), code_hits)


def test_cpu_tensors_are_rejected_loudly():
    import pytest, torch
    from cnns_slfp_quantization_b200.utils.sfp_quant import quantize_act, act_quantize_func
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        quantize_act(8)(torch.randn(8))
    assert quantize_act(32)(torch.ones(2)).tolist() == [1.0, 1.0]
    with pytest.raises(AssertionError):
        act_quantize_func(9)
    with pytest.raises(UnboundLocalError):
        act_quantize_func(5)(torch.ones(2))
