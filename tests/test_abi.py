"""The C-ABI library loads and exports every symbol include/salg.h declares (no compute, no GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "salg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(salg_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    import single_algebra_b200 as s
    lib = s._native.load()
    names = _declared()
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # and the ctypes prototypes cover the same set
    bound = set(s._native.PROTOTYPES) | set(s._native._SPECIAL)
    assert set(names) == bound, (set(names) ^ bound)


def test_version_and_error_string_without_gpu():
    import single_algebra_b200 as s
    lib = s._native.load()
    assert lib.salg_version() == 100
    assert lib.salg_prof_count() >= 10 and lib.salg_prof_name(0) == b"spmm"
    p = s._native.PcaParams()
    assert lib.salg_pca_params_default(ctypes.byref(p)) == 0
    assert (p.n_components, p.svd_method, p.center, p.random_seed) == (50, 0, 1, 42)
    assert lib.salg_pca_params_default(None) == 1 and b"NULL" in lib.salg_last_error()


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point fails loudly (SALG_ERR_CUDA)."""
    import single_algebra_b200 as s
    if s.device_count() > 0:
        return
    try:
        s.Context(0)
    except s.SalgError as e:
        assert e.code == 4 and "no CPU fallback" in str(e)
    else:
        raise AssertionError("context creation must fail without a GPU")


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "single-algebra_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
