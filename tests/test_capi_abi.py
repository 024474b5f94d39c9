"""CPU-only: the C-ABI library loads, exports every symbol include/pagmo_cuda/pgc.h declares, and reports errors
(not results) when there is no CUDA device - the product has no CPU fallback."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "pagmo_cuda" / "pgc.h"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(pgc_[a-z0-9_]+)\s*\(", text)))


def test_header_is_plain_c(tmp_path):
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "pagmo_cuda/pgc.h"\nint main(void){pgc_problem_desc d; (void)d; return 0;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", str(ROOT / "include"), "-c", str(src), "-o",
                    str(tmp_path / "t.o")], check=True)


def test_library_exports_every_declared_symbol():
    from pagmo2_b200 import capi
    L = capi.lib()
    syms = declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert b"sm_100a" in L.pgc_version()


def test_no_silent_cpu_fallback():
    """Without a device, context creation must fail with an error status (and a message), never compute."""
    from pagmo2_b200 import capi
    L = capi.lib()
    n = ctypes.c_int(-1)
    rc = L.pgc_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a CUDA device is visible here")
    h = ctypes.c_void_p()
    rc = L.pgc_ctx_create(0, ctypes.byref(h))
    assert rc in (capi.PGC_ERR_CUDA, capi.PGC_ERR_INVALID_ARGUMENT)
    assert not h.value
    assert L.pgc_last_error()
    with pytest.raises(capi.PgcError):
        capi.Context(0)


def test_product_does_not_reference_the_oracle():
    """The product tree must not include/link/import anything under oracle/."""
    bad = []
    for p in list((ROOT / "pagmo2_b200").rglob("*")) + list((ROOT / "include").rglob("*")):
        if p.is_file() and p.suffix in {".py", ".cu", ".cuh", ".cpp", ".h", ".hpp"}:
            t = p.read_text(errors="replace")
            if re.search(r'#include\s*[<"][^>"]*oracle|from\s+oracle|import\s+oracle|liboracle|libpagmo_ref', t):
                bad.append(str(p))
    assert not bad, bad
