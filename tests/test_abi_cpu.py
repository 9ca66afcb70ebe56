"""CPU tests of the boundary: the library loads without a GPU and exports every symbol the header declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from sessionsimilaritysearch_b200 import build
    return build.build()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sss_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sss_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    syms = header_symbols()
    for s in ["sss_index_create", "sss_index_add", "sss_index_search", "sss_normalize", "sss_topk_merge",
              "sss_binary_search", "sss_encoder_forward", "sss_item_vote", "sss_last_error"]:
        assert s in syms


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for s in header_symbols():
        assert hasattr(lib, s), "libsss_b200.so does not export " + s


def test_ctypes_binding_covers_header(built_lib):
    from sessionsimilaritysearch_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_symbols()
    lib = _lib.load()
    assert lib.sss_built_for_sm() == 100


def test_no_gpu_fails_loudly(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import sessionsimilaritysearch_b200 as sss
    with pytest.raises(RuntimeError):
        sss.IndexFlatIP(128)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "sessionsimilaritysearch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "libsss_oracle" not in txt, f


def test_built_sass_is_blackwell_native(built_lib):
    import subprocess
    out = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out or "SM100a" in out.upper().replace("_", "")
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in out, mnemonic + " missing from SASS"


def test_product_code_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under the package (Python or CUDA) may import, link or execute it, and
    there is no CPU fallback — the façade raises when the CUDA library is missing."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "sessionsimilaritysearch_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) in ("build", "__pycache__"):
            continue
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h")):
                continue
            text = open(os.path.join(dirpath, f), encoding="utf-8").read()
            # imports, includes, dlopen / subprocess of anything under oracle/ (mentions in comments are fine)
            if re.search(r"^\s*(from|import)\s+oracle\b|^\s*#\s*include\s*[<\"][^>\"]*oracle|dlopen\([^)]*oracle|"
                         r"(subprocess|os\.system)[^\n]*oracle", text, re.M):
                offenders.append(os.path.relpath(os.path.join(dirpath, f), root))
    assert offenders == [], offenders
    from sessionsimilaritysearch_b200 import _lib
    saved, _lib._lib, _lib.LIB_PATH = (_lib._lib, _lib.LIB_PATH), None, os.path.join(pkg, "no_such_library.so")
    try:
        import pytest
        with pytest.raises(RuntimeError, match="no\\s+CPU fallback"):
            _lib.load()
    finally:
        _lib._lib, _lib.LIB_PATH = saved
