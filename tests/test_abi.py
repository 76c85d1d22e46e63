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


def test_header_is_plain_c_and_links(tmp_path):
    """include/bgs.h is the drop-in boundary: it must compile as C99 (no C++isms) and a C program must link against
    libbgs.so and get a status code (BGS_ECUDA without a GPU) rather than a crash."""
    import subprocess, textwrap, torch

    root = os.path.join(os.path.dirname(__file__), "..")
    src = tmp_path / "abi_c.c"
    src.write_text(textwrap.dedent("""
        #include <stdio.h>
        #include "bgs.h"
        int main(void) {
            bgs_context* ctx = NULL;
            bgs_settings s = {BGS_GAUSSIAN_3D, BGS_RASTERIZE_COLOR, 0, 1, BGS_DRAW_ALL, 32, BGS_FLAG_ASYNC | BGS_FLAG_NO_CHUNKS, 0};
            bgs_cloud_uniform u = {{1,0,0,0, 0,1,0,0, 0,0,1,0, 0,0,0,1}, 1.0f, 1.0f, 0u, 0.0f, {0,0,0,1}, {1,1,1,1}};
            bgs_frame_stats fs;
            bgs_status st = bgs_context_create(0, &ctx);
            printf("%d %u %u %u\\n", (int)st, (unsigned)sizeof(s), (unsigned)sizeof(u), (unsigned)sizeof(fs));
            if (st == BGS_OK) bgs_context_destroy(ctx);
            return 0;
        }
    """))
    exe = tmp_path / "abi_c"
    libdir = os.path.join(root, "bevy_gaussian_splatting_b200")
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-lbgs", f"-Wl,-rpath,{libdir}"], check=True, capture_output=True, text=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out[1:] == ["32", "112", "40"]
    assert int(out[0]) == (abi.BGS_OK if torch.cuda.is_available() else abi.BGS_ECUDA)
