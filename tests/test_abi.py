"""The C-ABI shared library builds for sm_100a, loads, and exports every symbol include/b200sim.h declares.
No compute calls here (no GPU on the CPU box)."""

import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def built_lib() -> Path:
    import __graft_entry__ as g

    return g.build_cuda()


def test_header_symbols_are_exported(built_lib: Path) -> None:
    from ksim_b200 import binding

    header = (ROOT / "include" / "b200sim.h").read_text()
    declared = set(re.findall(r"\b(b200sim_[a-z_]+)\s*\(", header))
    assert declared == set(binding.EXPORTS)
    lib = binding.lib()
    for sym in declared:
        assert hasattr(lib, sym), sym
    nm = subprocess.run(["nm", "-D", "--defined-only", str(built_lib)], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (b200sim_[a-z_]+)", nm))
    assert declared <= exported


def test_library_contains_sm100a_code(built_lib: Path) -> None:
    out = subprocess.run(["cuobjdump", "-lelf", str(built_lib)], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_bad_arguments_return_error_codes(built_lib: Path) -> None:
    import ctypes

    from ksim_b200 import binding, codes as C

    lib = binding.lib()
    handle = ctypes.c_void_p()
    assert lib.b200sim_model_create(None, 0, ctypes.byref(handle)) == C.B2S_ERR_BAD_ARG
    junk = (ctypes.c_char * 16)()
    assert lib.b200sim_model_create(junk, 16, ctypes.byref(handle)) == C.B2S_ERR_BAD_BLOB
    assert b"header" in lib.b200sim_last_error()
    assert lib.b200sim_gae(None, 4, 4, None, None, None, None, 0.9, 0.9, 0, None, None, None) == C.B2S_ERR_BAD_ARG


def test_engine_refuses_cpu_device(walking_spec) -> None:  # noqa: ANN001
    from ksim_b200 import binding
    from ksim_b200.engine import B200Engine

    with pytest.raises(binding.B200SimError):
        B200Engine(walking_spec, 4, device="cpu")
