"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol that
include/lemas_b200.h declares; the ctypes mirrors match the C struct sizes; the product path fails loudly without a
device (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "lemas_b200.h").read_text()


def _declared_symbols():
    names = set(re.findall(r"\b(lemas_[a-z0-9_]+)\s*\(", HEADER))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    from lemas_tts import _native as nv

    lib = nv.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/lemas_b200.h but not exported"
    assert set(declared) == set(nv.SIGNATURES), "ctypes SIGNATURES and the header must list the same entry points"
    assert lib.lemas_version() >= 100


def test_ctypes_structs_match_c_layout():
    from lemas_tts import _native as nv

    lib = nv.load()
    for idx, cls in enumerate([nv.GemmDesc, nv.DitConfig, nv.DitLayer, nv.DitWeights, nv.SampleArgs, nv.VocosLayer,
                               nv.VocosWeights, nv.TextBlock, nv.TextWeights, nv.ProsodyTdnn, nv.ProsodyBlock,
                               nv.ProsodyWeights, nv.BigvganBlock, nv.BigvganStage, nv.BigvganWeights]):
        assert lib.lemas_abi_sizeof(idx) == C.sizeof(cls), cls.__name__


def test_no_device_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from lemas_tts import _native as nv, ops

    assert nv.load().lemas_device_supported() == 0
    with pytest.raises(RuntimeError, match="CUDA error"):
        nv.require_device()
    with pytest.raises(RuntimeError, match="CUDA error"):
        ops.ln_modulate(torch.zeros(8, 128), torch.zeros(128), torch.zeros(128))


def test_invalid_arguments_are_rejected_before_any_launch():
    from lemas_tts import _native as nv

    lib = nv.load()
    assert lib.lemas_gemm_f16(None, None) == 1  # LEMAS_ERR_INVALID
    assert b"null" in lib.lemas_last_error()
    d = nv.GemmDesc()
    assert lib.lemas_gemm_f16(C.byref(d), None) == 1
    cfg = nv.DitConfig(1024, 22, 16, 2, 512, 100, 16)
    assert lib.lemas_engine_workspace_bytes(C.byref(cfg), 1, 2187, 32) > 0
    assert lib.lemas_engine_workspace_bytes(None, 1, 1, 1) == -1


def test_missing_library_message(monkeypatch, tmp_path):
    from lemas_tts import _native as nv

    monkeypatch.setattr(nv, "_lib", None)
    monkeypatch.setenv("LEMAS_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="CUDA error.*missing"):
        nv.load()
