"""ctypes binding of liblemas_b200.so (C ABI in include/lemas_b200.h).

This is the stub INTEGRATION.md describes: torch tensors provide device memory and the current CUDA
stream; every FLOP of the hot path runs in the library's sm_100a kernels.  There is NO fallback: if the
library is missing or the device is not a Blackwell part, calls raise RuntimeError whose text contains
"CUDA error" / "no kernel image is available for execution on the device" — the substrings the reference
entry points match to trigger their own CPU retry (scripts/inference_gradio.py:317-321).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

_LIB_PATH = Path(__file__).resolve().parent.parent / "lib" / "liblemas_b200.so"
_lib = None

EPI_BIAS_F16, EPI_QKV_ROPE, EPI_GELU_TANH_F16, EPI_GELU_ERF_F16, EPI_GATE_RESID_F32 = 0, 1, 2, 3, 4
EPI_BIAS_F32, EPI_ADD_F32_F16, EPI_MISH_F16, EPI_MISH_RESID_F32 = 5, 6, 7, 8
SAMPLE_SKIP_PADDED_ROWS = 1
SAMPLE_FOLD_LAYERNORM = 2
PROF_KINDS = ["preloop", "in_proj", "conv_pos", "ln_mod", "gemm_qkv", "attention", "gemm_out", "gemm_ff1", "gemm_ff2",
              "proj_out", "cfg_euler", "tare"]

i32, i64, f32, vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class GemmDesc(C.Structure):
    _fields_ = [
        ("a", vp), ("batches", i32), ("rows", i32), ("lda", i32), ("a_cols", i32),
        ("w", vp), ("w_rows", i32), ("ldw", i32), ("n", i32), ("k_per_tap", i32),
        ("taps", i32), ("tap_pad", i32), ("w_tap_stride", i32), ("group_cols", i32),
        ("block_n", i32), ("epilogue", i32), ("bias", vp),
        ("out16", vp), ("ld16", i32), ("out32", vp), ("ld32", i32),
        ("resid", vp), ("ldr", i32), ("gate", vp), ("gate_bstride", i32),
        ("row_valid", vp), ("seq_len", i32), ("rope", vp), ("rope_cols", i32), ("inner", i32),
        ("vt", vp), ("vt_ld", i32), ("max_ctas", i32), ("row_limit", vp),
        ("ln_scale", vp), ("ln_out16", vp), ("ln_ld16", i32), ("ln_stats", vp),
        ("ln_stats_in", vp), ("ln_parts", i32), ("ln_uv", vp), ("ln_step", vp), ("ln_k", i32),
        ("tap_dilation", i32),
    ]


class DitConfig(C.Structure):
    _fields_ = [("dim", i32), ("depth", i32), ("heads", i32), ("ff_mult", i32), ("text_dim", i32),
                ("mel_dim", i32), ("rope_heads", i32)]


class DitLayer(C.Structure):
    _fields_ = [("w_qkv", vp), ("b_qkv", vp), ("w_out", vp), ("b_out", vp),
                ("w_ff1", vp), ("b_ff1", vp), ("w_ff2", vp), ("b_ff2", vp)]


class DitWeights(C.Structure):
    _fields_ = [("time_w0", vp), ("time_b0", vp), ("time_w2", vp), ("time_b2", vp),
                ("adaln_w", vp), ("adaln_b", vp), ("w_in_x", vp), ("w_in_ct", vp), ("b_in", vp), ("ct_ld", i32),
                ("conv_w", vp * 2), ("conv_b", vp * 2), ("conv_dense", i32),
                ("w_proj", vp), ("b_proj", vp), ("layers", C.POINTER(DitLayer))]


class SampleArgs(C.Structure):
    _fields_ = [("batch", i32), ("seq", i32), ("steps", i32), ("t_grid_host", C.POINTER(f32)),
                ("cfg_strength", f32), ("y", vp), ("step_cond", vp), ("text_cond", vp), ("text_uncond", vp),
                ("kv_len", vp), ("rope", vp), ("trajectory", vp), ("workspace", vp), ("workspace_bytes", i64),
                ("use_graph", i32), ("flags", i32),
                ("split_xchg_local", vp), ("split_xchg_peer", vp), ("split_flags_local", vp), ("split_flags_peer", vp),
                ("split_variant", i32)]


class VocosLayer(C.Structure):
    _fields_ = [("dw_w", vp), ("dw_b", vp), ("ln_w", vp), ("ln_b", vp), ("w1", vp), ("b1", vp),
                ("w2", vp), ("b2", vp), ("gamma", vp)]


class VocosWeights(C.Structure):
    _fields_ = [("dim", i32), ("inter", i32), ("layers", i32), ("in_ch", i32),
                ("embed_w", vp), ("embed_b", vp), ("norm_w", vp), ("norm_b", vp),
                ("blocks", C.POINTER(VocosLayer)), ("final_w", vp), ("final_b", vp),
                ("head_w", vp), ("head_b", vp)]


class TextBlock(C.Structure):
    _fields_ = [("dw_w", vp), ("dw_b", vp), ("ln_w", vp), ("ln_b", vp), ("w1", vp), ("b1", vp),
                ("grn_gamma", vp), ("grn_beta", vp), ("w2", vp), ("b2", vp)]


class TextWeights(C.Structure):
    _fields_ = [("dim", i32), ("inter", i32), ("layers", i32), ("mask_padding", i32), ("table", vp), ("pos", vp),
                ("blocks", C.POINTER(TextBlock))]


class ProsodyTdnn(C.Structure):
    _fields_ = [("w", vp), ("b", vp), ("ln_w", vp), ("ln_b", vp),
                ("cin", i32), ("cout", i32), ("k", i32), ("dil", i32), ("groups", i32)]


class ProsodyBlock(C.Structure):
    _fields_ = [("tdnn1", ProsodyTdnn), ("res2", ProsodyTdnn * 7), ("tdnn2", ProsodyTdnn),
                ("se_w1", vp), ("se_b1", vp), ("se_w2", vp), ("se_b2", vp)]


class ProsodyWeights(C.Structure):
    _fields_ = [("input_dim", i32), ("channels", i32), ("n_blocks", i32), ("scale", i32), ("se_channels", i32),
                ("att_channels", i32), ("mfa_channels", i32), ("embed_dim", i32),
                ("block0", ProsodyTdnn), ("blocks", C.POINTER(ProsodyBlock)), ("mfa", ProsodyTdnn),
                ("asp_tdnn", ProsodyTdnn), ("asp_conv_w", vp), ("asp_conv_b", vp), ("asp_norm_w", vp),
                ("asp_norm_b", vp), ("fc_w", vp), ("fc_b", vp)]


class BigvganBlock(C.Structure):
    _fields_ = [("kernel", i32), ("dilation", i32 * 3), ("w1", vp * 3), ("b1", vp * 3), ("w2", vp * 3), ("b2", vp * 3),
                ("act", vp * 6)]


class BigvganStage(C.Structure):
    _fields_ = [("rate", i32), ("ch_in", i32), ("ch_out", i32), ("up_w", vp), ("up_b", vp), ("block", BigvganBlock * 3)]


class BigvganWeights(C.Structure):
    _fields_ = [("num_mels", i32), ("ch0", i32), ("stages", i32), ("use_tanh", i32), ("pre_w", vp), ("pre_b", vp),
                ("stage", C.POINTER(BigvganStage)), ("post_act", vp), ("post_w", vp), ("post_bias", f32),
                ("aa_filter", f32 * 12)]


# every symbol include/lemas_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "lemas_last_error": (C.c_char_p, []),
    "lemas_version": (C.c_int, []),
    "lemas_device_supported": (C.c_int, []),
    "lemas_peer_alloc": (C.c_int, [i64, C.POINTER(C.c_void_p), vp]),
    "lemas_peer_open": (C.c_int, [vp, C.POINTER(C.c_void_p)]),
    "lemas_peer_close": (C.c_int, [vp]),
    "lemas_peer_free": (C.c_int, [vp]),
    "lemas_abi_sizeof": (C.c_int, [C.c_int]),
    "lemas_launch_count": (i64, []),
    "lemas_engine_profile": (C.c_int, [vp, i32]),
    "lemas_engine_profile_read": (C.c_int, [vp, C.POINTER(C.c_double), C.POINTER(i64), vp]),
    "lemas_gemm_f16": (C.c_int, [C.POINTER(GemmDesc), vp]),
    "lemas_ln_modulate": (C.c_int, [vp, vp, vp, i32, vp, i32, i32, i32, vp]),
    "lemas_ln_modulate_rows": (C.c_int, [vp, vp, vp, i32, vp, i32, i32, i32, vp, vp]),
    "lemas_ln_affine": (C.c_int, [vp, vp, vp, vp, vp, i32, i32, f32, vp]),
    "lemas_attention_f16": (C.c_int, [vp, i32, vp, i32, vp, vp, i32, i32, i32, vp]),
    "lemas_skinny_linear_f32": (C.c_int, [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "lemas_time_sinusoid": (C.c_int, [vp, vp, i32, vp]),
    "lemas_pack_cond_text": (C.c_int, [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "lemas_cast_pad_f16": (C.c_int, [vp, vp, i32, i32, i32, i32, vp]),
    "lemas_cfg_euler": (C.c_int, [vp, i32, vp, vp, i32, i32, vp, i32, i32, f32, f32, f32, vp]),
    "lemas_dwconv7_ln": (C.c_int, [vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    "lemas_istft_1024": (C.c_int, [vp, i32, vp, vp, i32, i32, vp]),
    "lemas_mel_spectrogram_1024": (C.c_int, [vp, i32, i32, i32, vp, vp, i32, vp, vp]),
    "lemas_mel_spectrogram_bigvgan_1024": (C.c_int, [vp, i32, i32, i32, vp, vp, i32, vp, vp]),
    "lemas_audio_prep_workspace_bytes": (i64, [i64]),
    "lemas_audio_prep": (C.c_int, [vp, i32, i64, i64, f32, vp, vp, vp, i64, vp]),
    "lemas_audio_unscale": (C.c_int, [vp, i64, vp, vp]),
    "lemas_audio_crossfade": (C.c_int, [vp, i64, vp, i64, i64, vp, C.c_double, vp]),
    "lemas_prosody_workspace_bytes": (i64, [C.POINTER(ProsodyWeights), i32, i32]),
    "lemas_prosody_encode": (C.c_int, [C.POINTER(ProsodyWeights), vp, i32, i32, vp, vp, i64, vp]),
    "lemas_resample_sinc": (C.c_int, [vp, i32, i32, i32, vp, i32, i32, i32, i32, vp, i32, i32, vp]),
    "lemas_kaldi_fbank_16k": (C.c_int, [vp, i32, i32, i32, vp, vp, vp, i32, vp, vp]),
    "lemas_engine_workspace_bytes": (i64, [C.POINTER(DitConfig), i32, i32, i32]),
    "lemas_engine_create": (C.c_int, [C.POINTER(DitConfig), C.POINTER(DitWeights), C.POINTER(vp)]),
    "lemas_engine_destroy": (None, [vp]),
    "lemas_sampler_run": (C.c_int, [vp, C.POINTER(SampleArgs), vp]),
    "lemas_dit_forward": (C.c_int, [vp, C.POINTER(SampleArgs), f32, vp, vp, vp]),
    "lemas_vocos_workspace_bytes": (i64, [C.POINTER(VocosWeights), i32, i32]),
    "lemas_vocos_decode": (C.c_int, [C.POINTER(VocosWeights), vp, vp, i32, i32, vp, i64, vp]),
    "lemas_bigvgan_workspace_bytes": (i64, [C.POINTER(BigvganWeights), i32, i32]),
    "lemas_bigvgan_decode": (C.c_int, [C.POINTER(BigvganWeights), vp, vp, i32, i32, vp, i64, vp]),
    "lemas_text_workspace_bytes": (i64, [C.POINTER(TextWeights), i32, i32]),
    "lemas_text_embedding": (C.c_int, [C.POINTER(TextWeights), vp, vp, vp, i32, i32, vp, i64, vp]),
}


def lib_path() -> Path:
    return Path(os.environ.get("LEMAS_B200_LIB", _LIB_PATH))


def load():
    """dlopen the library and type every entry point.  Raises loudly if the build is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not path.exists():
        raise RuntimeError(
            f"CUDA error: {path} is missing — build it with `python lemas-tts_b200/build.py` "
            "(this package has no CPU or PyTorch fallback)")
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().lemas_last_error().decode(errors="replace")
        raise RuntimeError(msg if "CUDA error" in msg else f"lemas_b200 error {rc}: {msg}")


def ptr(t: torch.Tensor | None):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "native ops take contiguous CUDA tensors"
    return t.data_ptr()


def stream() -> int:
    """Raw handle of the CURRENT device's current stream — call it inside `on_device` (below)."""
    return torch.cuda.current_stream().cuda_stream


def on_device(fn):
    """Run a native call on the device that owns its operands: the first CUDA tensor among the arguments, else the
    `.device` of `self`.  The library launches on the current device with the stream it is handed, and the reference
    API accepts any device string (`TTS(device="cuda:1")`, api.py:94) without a prior torch.cuda.set_device — so the
    wrapper, not the caller, selects the device (and thereby the stream `stream()` returns)."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = None
        if args and not isinstance(args[0], torch.Tensor):  # engine method: the engine's own device wins
            d = getattr(args[0], "device", None)
            if isinstance(d, torch.device) and d.type == "cuda":
                dev = d
        if dev is None:
            for a in list(args) + list(kwargs.values()):
                if isinstance(a, torch.Tensor) and a.is_cuda:
                    dev = a.device
                    break
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)

    return wrapper


def require_device() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("CUDA error: no CUDA device is available (lemas_tts B200 build has no CPU path)")
    if not load().lemas_device_supported():
        raise RuntimeError("CUDA error: no kernel image is available for execution on the device "
                           "(liblemas_b200 is built for sm_100a only)")


def weights_key(module) -> tuple:
    """Cache key of the packed (fp16 / re-laid-out) copy of a module's weights: device, storage addresses and the
    in-place version counters of EVERY parameter and buffer — an EMA `p.data.copy_`, a partial load or a
    `torch.nn.utils` helper on any tensor invalidates the packed copy, not only a change of one sentinel weight.
    Tensors created under torch.inference_mode() have no version counter; their address still takes part."""
    ver, ptrs, n, dev = 0, 0, 0, None
    for t in list(module.parameters()) + list(module.buffers()):
        try:
            ver += t._version
        except RuntimeError:
            pass
        ptrs = (ptrs * 1000003 + t.data_ptr()) & 0xFFFFFFFFFFFFFFFF
        n += 1
        dev = t.device
    return (str(dev), n, ver, ptrs)


def state_fingerprint(sd: dict) -> str:
    """Content fingerprint of a state dict (names, shapes, dtypes, fp64 sum and a strided sample sum of every tensor):
    the file name of the packed-weight blob that belongs to these weights (LEMAS_PACKED_CACHE)."""
    import hashlib

    h = hashlib.sha1()
    sums = []
    for k in sorted(sd):
        v = sd[k]
        h.update(f"{k}|{tuple(v.shape)}|{v.dtype}|".encode())
        f = v.detach().reshape(-1)
        sums.append(torch.stack((f.double().sum(), f[:: 997].double().sum())))
    if sums:
        h.update(torch.stack(sums).cpu().numpy().tobytes())
    return h.hexdigest()[:20]
