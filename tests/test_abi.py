"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares
(no compute calls without a GPU), and the product path never touches the oracle."""
import ctypes
import os
import re
import subprocess

from conftest import ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "dcb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dcb_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from dynamicrafter_b200 import build, _lib
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _header_symbols()
    assert len(declared) >= 19
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in dcb200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared
    assert lib.dcb_abi_version() == 1


def test_argument_validation_without_gpu():
    from dynamicrafter_b200 import _lib
    L = _lib.lib()
    a = _lib.TapGemmArgs()
    assert L.dcb_tap_gemm(ctypes.byref(a), None) == -1          # DCB_ERR_INVALID, no launch attempted
    assert b"null" in L.dcb_last_error()
    assert L.dcb_layer_norm(None, 1, 64, None, None, ctypes.c_float(1e-5), None, None) == -1


def test_struct_layout_matches_header():
    """sizeof of the ctypes mirrors equals what the C compiler lays out for include/dcb200.h."""
    from dynamicrafter_b200 import _lib
    prog = r'''
#include <stdio.h>
#include "dcb200.h"
int main(void){ printf("%zu %zu %zu\n", sizeof(dcb_tap_gemm_args), sizeof(dcb_attention_args), sizeof(dcb_ddim_coeffs)); return 0; }
'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(_lib.TapGemmArgs), ctypes.sizeof(_lib.AttentionArgs), ctypes.sizeof(_lib.DdimCoeffs)]


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dynamicrafter_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(dp, f)


def test_cpu_tensors_fail_loudly():
    import pytest
    import torch
    from dynamicrafter_b200 import ops, _lib
    with pytest.raises(_lib.DcbError):
        ops.layer_norm(torch.zeros(4, 64), torch.ones(64), torch.zeros(64))
