"""The C++ host facade (include/single_algebra.hpp) — the compiled-language mirror of the reference's Rust type surface —
driven by the reference's own tests restated in C++ (tests/cpp/facade_test.cpp: csr.rs:1514-1550,
pca/sparse/mod.rs:540-562, the masked type's error strings)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "facade_test.cpp")
OUT_DIR = os.path.join(ROOT, "tests", "cpp", "build")
EXE = os.path.join(OUT_DIR, "facade_test")


def build_facade_test():
    os.makedirs(OUT_DIR, exist_ok=True)
    lib_dir = os.path.join(ROOT, "single-algebra_b200")
    assert os.path.exists(os.path.join(lib_dir, "libsalg_b200.so")), "build the CUDA extension first (__graft_entry__.build)"
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC,
           "-L", lib_dir, "-lsalg_b200", "-Wl,-rpath," + lib_dir, "-o", EXE]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return EXE


def test_facade_compiles_and_refuses_to_run_without_a_gpu():
    exe = build_facade_test()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    import single_algebra_b200 as s
    if s.device_count() == 0:
        # no CPU fallback: the program must say so and stop
        assert r.returncode == 2 and "no CPU fallback" in r.stdout, (r.returncode, r.stdout, r.stderr)
    else:
        assert r.returncode == 0, (r.stdout, r.stderr)


@pytest.mark.gpu
def test_reference_tests_through_the_cpp_facade():
    exe = build_facade_test()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "all facade tests passed" in r.stdout, (r.stdout[-3000:], r.stderr[-2000:])
    for name in ("test_csr_normalize", "test_sums_and_log1p<f32>", "test_sums_and_log1p<f64>", "test_csc_twins",
                 "test_random_matrix_sparse_svd_comp_random", "test_masked"):
        assert "ok " + name in r.stdout, r.stdout
