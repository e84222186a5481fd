"""CPU-side checks of the C-ABI library: it builds for sm_100a, loads, exports every
symbol include/ipfa_b200.h declares, and the product path refuses to run without CUDA
(no CPU fallback, no route through oracle/)."""
import ast
import glob
import importlib
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "iterative-pseudo-forced-alignment-ctc_b200"


@pytest.fixture(scope="module")
def libmod():
    return importlib.import_module(PKG + "._lib")


def test_library_exports_every_declared_symbol(libmod):
    L = libmod.lib()
    declared = libmod.declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(L, name), name
        assert name in libmod._SIGNATURES, f"{name} has no ctypes signature"
    assert L.ipfa_version() >= 100
    assert L.ipfa_status_string(0) == b"ok"
    assert b"shorter than text" in L.ipfa_status_string(5)


def test_sass_is_sm100a_only(libmod):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", libmod._build.LIB], capture_output=True, text=True).stdout
    archs = {ln.split(".")[-2] for ln in out.splitlines() if "sm_" in ln}
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback():
    import torch
    ipfa = importlib.import_module(PKG)
    lp = torch.zeros(1, 4, 3).log_softmax(-1)
    if not torch.cuda.is_available():
        with pytest.raises(ValueError, match="CUDA tensor"):
            ipfa.ctc_alpha_nll(lp, [[1]], [4], [1])
        with pytest.raises(ValueError, match="CUDA tensor"):
            ipfa.ctc_forced_align(lp, [[1]], [4], [1])
        # host-buffer entry point: fails loudly with the CUDA error, does not compute on the CPU
        with pytest.raises(RuntimeError, match="CUDA"):
            ipfa.ctc_alpha_nll_host(lp.numpy(), np.array([[1]], np.int32), [4], [1])


def test_product_package_never_imports_oracle():
    for path in glob.glob(os.path.join(ROOT, PKG, "**", "*.py"), recursive=True) + \
            glob.glob(os.path.join(ROOT, "src", "*.py")):
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            assert not any(n == "oracle" or n.startswith("oracle.") for n in names), path


def test_invalid_arguments_are_rejected(libmod):
    L = libmod.lib()
    assert L.ipfa_ctc_alpha_device(None, 0, 0, None, 0, None, None, 4, 10, 2, 8, 0, None, None, 0, None) == 1
    assert L.ipfa_ctc_alpha_workspace_bytes(4, 10, 2, 8) > 0
    assert L.ipfa_ctc_viterbi_workspace_bytes(4, 10, 2, 8) >= 4 * 3 * 32 * 4
    assert L.ipfa_ctcseg_workspace_bytes(2, 100, 20, 3, 8) > 0


def test_alpha_workspace_holds_its_documented_parts(libmod):
    """include/ipfa_b200.h: N + 4 counter words, the two halves' state vectors of every window
    ([N][2][2][Lmax + 1] fp32), the two length-bucket lists and the two tiers' redo lists ([4][N] int32)."""
    L = libmod.lib()
    for n, lmax in ((1, 0), (7, 1), (1024, 100), (4300, 40), (65536, 40)):
        need = (n + 4) * 4 + n * 4 * (lmax + 1) * 4 + 4 * n * 4
        got = L.ipfa_ctc_alpha_workspace_bytes(n, 1000, lmax, 32)
        assert need <= got <= need + 8 * 256, (n, lmax, got, need)
    # monotone in N (the host entry point sizes one workspace for all its chunks)
    sizes = [L.ipfa_ctc_alpha_workspace_bytes(n, 1000, 100, 32) for n in (1, 100, 132, 1024)]
    assert sizes == sorted(sizes)
