import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Every test process needs libsagars.so (even CPU tests check its ABI): build it if it is missing."""
    from seganygaussians_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    yield


@pytest.fixture(scope="session", autouse=True)
def _emulated_kernels_under_asan():
    """SAGARS_EMU_ASAN=1 LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:verify_asan_link_order=0
    python -m pytest tests -m "not gpu" -k emulated   -- every kernel that the CPU execution shim (tests/cuda_emu) runs is built with
    AddressSanitizer: an out-of-bounds access to a global array or to the block's shared-memory buffer aborts the run with the
    kernel's source line (the CPU-side counterpart of `compute-sanitizer --tool memcheck`, profiles/r2_sanitizer.md)."""
    san = "address" if os.environ.get("SAGARS_EMU_ASAN") == "1" else ("thread" if os.environ.get("SAGARS_EMU_TSAN") == "1" else None)
    if san is None:
        yield
        return
    import subprocess
    orig = subprocess.check_call

    def check_call(cmd, *a, **kw):
        if isinstance(cmd, (list, tuple)) and cmd and cmd[0] == "g++" and "-shared" in cmd and any("cuda_emu" in str(c) for c in cmd):
            cmd = [cmd[0], "-fsanitize=" + san, "-fno-omit-frame-pointer", "-g"] + list(cmd[1:])
        return orig(cmd, *a, **kw)

    subprocess.check_call = check_call
    try:
        yield
    finally:
        subprocess.check_call = orig
