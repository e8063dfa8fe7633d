"""CPU tier: the C-ABI library loads, exports exactly the symbols include/foundation_pt.h declares, the ctypes mirror of
the structs matches the header's sizes, and — with no GPU — the product fails loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess

import pytest

from foundation_b200 import build, pt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "foundation_pt.h")).read()
    return sorted(set(re.findall(r"FOUNDATION_PT_API\s+[\w\s\*]+?\b(foundation_pt_\w+)\s*\(", src)))


def test_library_builds_for_sm_100a_and_exports_every_declared_symbol():
    lib_path = build.build()
    assert os.path.exists(lib_path)
    declared = header_symbols()
    assert len(declared) >= 24 and set(declared) == set(pt.SYMBOLS), set(declared) ^ set(pt.SYMBOLS)
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r"\bT\s+(foundation_pt_\w+)", out)))
    assert exported == declared, set(exported) ^ set(declared)
    lib = pt.load_library()
    assert lib.foundation_pt_version() == 0x00010000


def test_cubin_is_sm_100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", build.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_(?!100a)\d+", out), out


def test_struct_sizes_match_header():
    src = f'#include <stdio.h>\n#include "{ROOT}/include/foundation_pt.h"\nint main(){{printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(foundation_pt_config), ' \
          'sizeof(foundation_pt_build_stats), sizeof(foundation_pt_stats), sizeof(foundation_pt_ray), sizeof(foundation_pt_hit), sizeof(foundation_pt_instance), ' \
          'sizeof(foundation_pt_material));return 0;}'
    exe = "/tmp/pt_sizes_test"
    subprocess.run(["gcc", "-std=c99", "-x", "c", "-", "-o", exe], input=src, text=True, check=True)      # the header is plain C99
    sizes = list(map(int, subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()))
    assert sizes == [C.sizeof(pt.Config), C.sizeof(pt.BuildStats), C.sizeof(pt.Stats), 32, 16, 64, 32]


def test_no_gpu_means_loud_failure_not_a_cpu_fallback():
    lib = pt.load_library()
    cfg = pt.Config(C.sizeof(pt.Config), 0, 16, 16, 1, 0, 0, (C.c_float * 3)(0, 0, 0), 0)
    ctx = C.c_void_p()
    st = lib.foundation_pt_create(C.byref(cfg), None, C.byref(ctx))
    if st == pt.OK:                      # running on a GPU box: nothing to check here
        lib.foundation_pt_destroy(ctx)
        pytest.skip("CUDA device present")
    assert st == pt.ERR_NO_DEVICE and not ctx.value
    assert b"no CPU fallback" in lib.foundation_pt_last_error(None)
    with pytest.raises(pt.FoundationPtError):
        pt.PathTracer(16, 16)
    # bad arguments are rejected before any device work
    bad = pt.Config(4, 0, 16, 16, 1, 0, 0, (C.c_float * 3)(0, 0, 0), 0)
    assert lib.foundation_pt_create(C.byref(bad), None, C.byref(ctx)) == pt.ERR_ARGUMENT
    assert lib.foundation_pt_destroy(None) == pt.ERR_ARGUMENT


def test_product_never_touches_the_oracle():
    """The product path must not import, include or link anything under oracle/ (tier rule ③)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "foundation_b200")):
        for f in files:
            if f.endswith((".py", ".h", ".cuh", ".cu", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r'(#include\s*"[^"]*oracle/|from\s+oracle|import\s+oracle|libpt_oracle|pt_emu)', txt), os.path.join(dirpath, f)


def test_cpp_renderer_host_links_against_the_c_abi_only():
    """The Foundation-shaped C++ host (Renderer(device, allocator) / Draw()) compiles with g++ alone and links only
    libfoundation_pt.so — no CUDA headers, no torch."""
    rdir = os.path.join(ROOT, "foundation_b200", "renderer")
    subprocess.run(["make", "-C", rdir, "-B"], check=True, capture_output=True)
    exe = os.path.join(rdir, "foundation_editor")
    assert os.path.exists(exe)
    needed = subprocess.run(["readelf", "-d", exe], capture_output=True, text=True).stdout
    assert "libfoundation_pt.so" in needed and "libcuda" not in needed and "libtorch" not in needed
    for f in ("Renderer.hpp", "Renderer.cpp", "Editor.cpp"):
        txt = open(os.path.join(rdir, f)).read()
        assert "cuda_runtime" not in txt and "torch" not in txt.replace("no torch", "")


def test_rhi_present_path_compiles_and_follows_the_reference_upload_protocol():
    """Present.cpp (abstract-RHI calls only) builds against the stub subset of the reference's RHI headers and,
    run against a recording mock, performs the staging upload of mos9527/Foundation src/Renderer/Renderer.cpp:219-251 with the resolved bytes."""
    rdir = os.path.join(ROOT, "foundation_b200", "renderer")
    subprocess.run(["make", "-C", rdir, "present_selftest"], check=True, capture_output=True)
    out = subprocess.run([os.path.join(rdir, "present_selftest")], capture_output=True, text=True)
    assert out.returncode == 0 and "present path OK" in out.stdout, out.stdout + out.stderr
