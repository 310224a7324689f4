"""CPU: the C-ABI library loads and exports every function include/stgconv_b200.h declares, the
ctypes mirrors agree with the header's struct layouts, and compute entry points fail loudly (status
code + message) instead of falling back when there is no GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "stgconv_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(stg_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_expected_surface():
    names = _declared_functions()
    for must in ("stg_block_forward", "stg_block_backward", "stg_block_xmoments", "stg_model_forward",
                 "stg_model_backward", "stg_model_loss_backward", "stg_adam_step", "stg_model_workspace_bytes",
                 "stg_last_error", "stg_version"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from gnn_rul_benchmarking_b200 import _lib
    lib = _lib.load()
    for name in _declared_functions():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.SIGNATURES"
    assert b"sm_100a" in lib.stg_version()


def test_struct_sizes_match_c_layout(tmp_path):
    """Compile a tiny C program against the header and compare sizeof/offsetof with ctypes."""
    import subprocess
    from gnn_rul_benchmarking_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "stgconv_b200.h"\n'
                   "int main(void){printf(\"%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(stg_block_desc),"
                   "sizeof(stg_block_grads), sizeof(stg_model_dims), sizeof(stg_bn), sizeof(stg_model_block),"
                   "sizeof(stg_model_params), sizeof(stg_dropout), offsetof(stg_model_params, blk),"
                   "offsetof(stg_model_params, fc_w));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(_lib.StgBlockDesc), C.sizeof(_lib.StgBlockGrads), C.sizeof(_lib.StgModelDims),
            C.sizeof(_lib.StgBN), C.sizeof(_lib.StgModelBlock), C.sizeof(_lib.StgModelParams),
            C.sizeof(_lib.StgDropout), _lib.StgModelParams.blk.offset, _lib.StgModelParams.fc_w.offset]
    assert got == want


def test_stats_macro_matches_python_mirror(tmp_path):
    import subprocess
    from gnn_rul_benchmarking_b200.functional import stats_doubles
    cases = [(16, 8, 25), (14, 7, 50), (48, 24, 50), (4, 2, 7), (6, 3, 9), (32, 16, 13)]
    body = "".join(f'printf("%d\\n", (int)STG_BLOCK_STATS_DOUBLES({c},{h},{t}));' for c, h, t in cases)
    src = tmp_path / "m.c"
    src.write_text(f'#include <stdio.h>\n#include "stgconv_b200.h"\nint main(void){{{body}return 0;}}\n')
    exe = tmp_path / "m"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [stats_doubles(c, h, t) for c, h, t in cases]


def test_invalid_arguments_return_status_not_crash():
    """Argument validation happens before any CUDA call, so it is testable without a GPU."""
    from gnn_rul_benchmarking_b200 import _lib
    lib = _lib.load()
    assert lib.stg_block_xmoments(None, 1, 1, 1, 1, None, None) == -1
    assert b"bad argument" in lib.stg_last_error()
    d = _lib.StgModelDims()
    assert lib.stg_model_workspace_bytes(C.byref(d)) == 0          # all-zero dims are invalid
    d.B, d.N, d.T, d.P, d.K, d.EH, d.E, d.H = 256, 14, 25, 2, 2, 8, 6, 8
    d.w[0] = d.w[1] = 2
    d.stride[0], d.stride[1] = 1, 2
    n = lib.stg_model_workspace_bytes(C.byref(d))
    assert n > 256 * 25 * 14 * 16 * 4 * 4                           # h, dh, two dxp at least
    rc = lib.stg_model_forward(C.byref(d), None, None, None, 0, 0, None, None, None)
    assert rc == -1
    with pytest.raises(ValueError):
        _lib.check(rc, "stg_model_forward")
    # sibling primitives: null pointers / impossible sizes are status codes, unsupported tiles say so
    assert lib.stg_patch_stats(None, 4, 16, None, None) == -1
    assert lib.stg_patch_stats11(None, 4, 16, None, None) == -1
    assert lib.stg_patch_stats12(None, 4, 16, None, None) == -1
    assert lib.stg_gat_forward(None, None, None, None, 0, None, 0.0, 0.1, 0.01, 2, 8, 16, None, None) == -1
    fake = C.c_void_p(256)                                           # never dereferenced: validation comes first
    assert lib.stg_gat_forward(fake, fake, fake, fake, 0, None, 0.0, 0.1, 0.0, 2, 8, 16, fake, None) == -1      # slope 0
    assert lib.stg_gat_forward(fake, fake, fake, fake, 0, None, 0.0, 0.1, 0.01, 2, 400, 400, fake, None) == -2  # tile
    assert b"does not fit" in lib.stg_last_error()
    assert lib.stg_adj_forward(9, fake, 2, 8, 4, 0, fake, None, None) == -1                                      # kind
    assert lib.stg_agg_forward(0, fake, fake, 2, 400, 64, fake, None) == -2


def test_product_package_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "gnn_rul_benchmarking_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
