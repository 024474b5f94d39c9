"""The C++ drop-in adapters (include/pagmo_cuda/cuda_bfe.hpp) behind pagmo's own type-erased problem / bfe /
population classes.  tests/cpp/test_adapters.cpp is compiled in the authoring container against the unmodified
reference headers (tests/cpp/Makefile, run by __graft_entry__.build()); the binary travels to the GPU box."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "tests" / "cpp" / "_bin" / "test_adapters"
BIN_ISLANDS = ROOT / "tests" / "cpp" / "_bin" / "test_islands"


def _ensure_binary():
    if not BIN.exists() or not BIN_ISLANDS.exists():
        if Path("/root/reference/include/pagmo/problem.hpp").exists():
            subprocess.run(["make", "-s", "-C", str(ROOT / "tests" / "cpp")], check=True)
        else:
            pytest.skip("tests/cpp/_bin/test_adapters was not prebuilt and /root/reference is absent")


def test_adapters_fail_loudly_without_a_device():
    """CPU box: the adapter must raise (std::runtime_error from pgc_ctx_create), not compute on the host."""
    from pagmo2_b200 import capi
    import ctypes
    n = ctypes.c_int(0)
    if capi.lib().pgc_device_count(ctypes.byref(n)) == 0 and n.value > 0:
        pytest.skip("a CUDA device is visible here")
    _ensure_binary()
    r = subprocess.run([str(BIN)], capture_output=True, text=True)
    assert r.returncode != 0
    assert "pgc_ctx_create" in (r.stderr + r.stdout)
    assert "ADAPTERS OK" not in r.stdout


@pytest.mark.gpu
def test_adapters_on_device():
    _ensure_binary()
    r = subprocess.run([str(BIN)], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ADAPTERS OK" in r.stdout


def test_islands_fail_loudly_without_a_device():
    from pagmo2_b200 import capi
    import ctypes
    n = ctypes.c_int(0)
    if capi.lib().pgc_device_count(ctypes.byref(n)) == 0 and n.value > 0:
        pytest.skip("a CUDA device is visible here")
    _ensure_binary()
    r = subprocess.run([str(BIN_ISLANDS)], capture_output=True, text=True)
    assert r.returncode != 0 and "ISLANDS OK" not in r.stdout
    assert "pgc_ctx_create" in (r.stderr + r.stdout)


@pytest.mark.gpu
def test_islands_on_device():
    """cuda_island inside pagmo's own island / archipelago{ring} (the reference's island.cpp, archipelago.cpp, ring.cpp compiled
    unmodified into oracle/_ref) and the device-resident cuda_archipelago; with two or more GPUs visible the migrants travel over
    NCCL and the run must equal the one-GPU run bit for bit."""
    import torch
    _ensure_binary()
    ndev = min(torch.cuda.device_count(), 8)
    r = subprocess.run([str(BIN_ISLANDS), "--devices", str(ndev)], capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ISLANDS OK" in r.stdout
