"""Drop-in boundary: libbgs.so loads, exports every symbol include/bgs.h declares, and refuses to
run without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from bevy_gaussian_splatting_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "bgs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bgs_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = abi.load()
    declared = _header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"libbgs.so does not export {name}"
    assert sorted(n for n, _, _ in abi.SYMBOLS) == declared, "abi.SYMBOLS and include/bgs.h disagree"


def test_struct_layouts_match_header():
    # bgs_view: 3*16 + 3 + 4 floats; bgs_cloud_uniform: 16 floats + 2 floats + u32 + float + 2 x vec4; bgs_settings: 8 u32
    assert C.sizeof(abi.bgs_view) == (48 + 3 + 4) * 4
    assert C.sizeof(abi.bgs_cloud_uniform) == 28 * 4 and abi.bgs_cloud_uniform.aabb_min.offset == 80
    assert C.sizeof(abi.bgs_settings) == 32
    assert C.sizeof(abi.bgs_frame_stats) == 40 and abi.bgs_frame_stats.rounds.offset == 32
    assert abi.bgs_frame_stats.n_pairs.offset == 8


def test_null_arguments_return_status_codes_not_crashes():
    lib = abi.load()
    assert lib.bgs_context_create(0, None) == abi.BGS_EINVAL
    assert lib.bgs_render(None, None, None, None, None, None, 0, 0) == abi.BGS_EINVAL
    assert lib.bgs_debug_sorted_entries(None, None) == abi.BGS_EINVAL
    assert lib.bgs_gather_frames(None, None, 0, None, None, 0) == abi.BGS_EINVAL
    assert lib.bgs_last_error(None) == b"null context"
    lib.bgs_context_destroy(None)
    lib.bgs_cloud_destroy(None)


def test_no_cpu_fallback_without_a_gpu():
    """On a box without a CUDA device the product must fail loudly, never compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the failure path is exercised on the CPU-only box")
    lib = abi.load()
    ctx = C.c_void_p()
    assert lib.bgs_context_create(0, C.byref(ctx)) == abi.BGS_ECUDA
    assert not ctx.value
    import bevy_gaussian_splatting_b200 as B

    with pytest.raises(abi.BgsError):
        B.GaussianSplattingPlugin(0)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may import, load or link it."""
    pkg = os.path.join(ROOT, "bevy_gaussian_splatting_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for line in text.splitlines():
                    code = line.split("//")[0].split("#")[0] if not f.endswith(".py") else line.split("#")[0]
                    assert "liboracle" not in code and "bgs_oracle" not in code, (f, line)
                    if f.endswith(".py"):
                        assert not re.search(r"^\s*(from|import)\s+oracle\b", code), (f, line)
