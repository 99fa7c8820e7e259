"""Host side of the native engine: packs reference-layout weights into the buffers liblemas_b200.so reads
and drives `lemas_sampler_run` / `lemas_dit_forward` / `lemas_vocos_decode` (include/lemas_b200.h).

Pure plumbing: torch owns the device memory and the stream, the library does every FLOP.  There is no fallback:
a missing library or a non-Blackwell device raises RuntimeError("CUDA error: ...") from `_native`.

Weight layout (reference checkpoint keys -> engine buffers):
  transformer.time_embed.time_mlp.{0,2}                       -> fp32 time_w0/time_w2 (pre-loop skinny GEMMs)
  transformer_blocks.{i}.attn_norm.linear, norm_out.linear    -> ONE stacked fp32 [depth*6*D + 2*D, D] matrix:
                                                                 all AdaLN modulations of all steps are produced
                                                                 before the ODE loop (they depend on t only)
  input_embed.proj  [D, mel | mel | text]                     -> fp16 x-columns [D,128] + fp16 (cond|text) columns
  input_embed.conv_pos_embed.conv1d.{0,2}  [D, D/16, 31]      -> fp16 tap-major [31*D, 64] (grouped, D/16 == 64) or
                                                                 block-diagonal dense [31*D, D]
  attn.to_q | to_k | to_v                                     -> fp16 [3*inner, D] stacked, fp32 bias
  attn.to_out.0, ff.ff.0.0, ff.ff.2, proj_out                 -> fp16, fp32 bias
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _native as nv

f16, f32 = torch.float16, torch.float32


def _dev(t: torch.Tensor, device, dtype) -> torch.Tensor:
    return t.detach().to(device=device, dtype=dtype).contiguous()


class DiTEngine:
    """Owns the packed DiT weights on one device and a native engine handle."""

    BLOB_VERSION = 1

    def __init__(self, sd: dict | None, *, dim: int, depth: int, heads: int, ff_mult: int, text_dim: int, mel_dim: int,
                 pe_attn_head: int | None = None, qk_norm: str | None = None, device="cuda",
                 prefix: str = "transformer.", packed: dict | None = None):
        """`sd`: state dict in the reference key layout (packed here: fp16 K-major GEMM operands, concatenated QKV /
        AdaLN matrices, tap-major conv weights), or `packed`: the tensors a previous `export_blob` wrote."""
        nv.require_device()
        if qk_norm is not None:
            raise RuntimeError("lemas_b200 error: qk_norm is not supported by the sm_100a engine "
                               "(both shipped configs set qk_norm: null)")
        self.device = torch.device(device)
        self.dim, self.depth, self.heads, self.ff_mult = dim, depth, heads, ff_mult
        self.text_dim, self.mel_dim = text_dim, mel_dim
        self.inner = heads * 64
        self.rope_heads = heads if pe_attn_head is None else int(pe_attn_head)
        self.pe_attn_head = pe_attn_head
        D, M, Dt, dv = dim, mel_dim, text_dim, self.device
        p = prefix
        self._packed = {}  # name -> tensor: everything the native side holds a pointer to (also what export_blob writes)

        def take(name, make):
            t = packed[name].to(dv) if packed is not None else make()
            self._packed[name] = t
            return nv.ptr(t)

        w = nv.DitWeights()
        w.time_w0 = take("time_w0", lambda: _dev(sd[p + "time_embed.time_mlp.0.weight"], dv, f32))
        w.time_b0 = take("time_b0", lambda: _dev(sd[p + "time_embed.time_mlp.0.bias"], dv, f32))
        w.time_w2 = take("time_w2", lambda: _dev(sd[p + "time_embed.time_mlp.2.weight"], dv, f32))
        w.time_b2 = take("time_b2", lambda: _dev(sd[p + "time_embed.time_mlp.2.bias"], dv, f32))

        def adaln(kind):
            parts = [sd[f"{p}transformer_blocks.{i}.attn_norm.linear.{kind}"] for i in range(depth)]
            parts.append(sd[p + f"norm_out.linear.{kind}"])
            return _dev(torch.cat([t.float() for t in parts], 0), dv, f32)

        w.adaln_w = take("adaln_w", lambda: adaln("weight"))
        w.adaln_b = take("adaln_b", lambda: adaln("bias"))

        self.ct_ld = (M + Dt + 63) // 64 * 64

        def in_proj(which):  # input_embed.proj.weight [D, 2*mel + text_dim]: columns x | cond | text
            proj = sd[p + "input_embed.proj.weight"].float()
            assert proj.shape == (D, 2 * M + Dt), proj.shape
            if which == "x":
                w_x = torch.zeros(D, 128)
                w_x[:, :M] = proj[:, :M]
                return _dev(w_x, dv, f16)
            w_ct = torch.zeros(D, self.ct_ld)
            w_ct[:, : M + Dt] = proj[:, M:]
            return _dev(w_ct, dv, f16)

        w.w_in_x = take("w_in_x", lambda: in_proj("x"))
        w.w_in_ct = take("w_in_ct", lambda: in_proj("ct"))
        w.b_in = take("b_in", lambda: _dev(sd[p + "input_embed.proj.bias"], dv, f32))
        w.ct_ld = self.ct_ld

        groups = 16
        gc = D // groups
        w.conv_dense = 0 if gc == 64 else 1

        def conv_weight(idx):
            cw = sd[f"{p}input_embed.conv_pos_embed.conv1d.{idx}.weight"].float()  # [D, D/16, 31]
            taps = cw.shape[-1]
            assert cw.shape == (D, gc, 31), cw.shape
            if gc == 64:
                return _dev(cw.permute(2, 0, 1).reshape(taps * D, gc), dv, f16)
            dense = torch.zeros(taps, D, D)  # block-diagonal dense: out channel o reads its own group's inputs only
            for g in range(groups):
                dense[:, g * gc:(g + 1) * gc, g * gc:(g + 1) * gc] = cw[g * gc:(g + 1) * gc].permute(2, 0, 1)
            return _dev(dense.reshape(taps * D, D), dv, f16)

        for j, idx in enumerate((0, 2)):
            w.conv_w[j] = take(f"conv_w{j}", lambda idx=idx: conv_weight(idx))
            w.conv_b[j] = take(f"conv_b{j}",
                               lambda idx=idx: _dev(sd[f"{p}input_embed.conv_pos_embed.conv1d.{idx}.bias"], dv, f32))

        def proj_out():
            w_proj = torch.zeros(128, D)
            w_proj[:M] = sd[p + "proj_out.weight"].float()
            return _dev(w_proj, dv, f16)

        w.w_proj = take("w_proj", proj_out)
        w.b_proj = take("b_proj", lambda: _dev(sd[p + "proj_out.bias"], dv, f32))

        layers = (nv.DitLayer * depth)()
        for i in range(depth):
            q = f"{p}transformer_blocks.{i}."
            L = layers[i]
            L.w_qkv = take(f"l{i}.w_qkv", lambda q=q: _dev(
                torch.cat([sd[q + f"attn.to_{n}.weight"].float() for n in "qkv"], 0), dv, f16))
            L.b_qkv = take(f"l{i}.b_qkv", lambda q=q: _dev(
                torch.cat([sd[q + f"attn.to_{n}.bias"].float() for n in "qkv"], 0), dv, f32))
            L.w_out = take(f"l{i}.w_out", lambda q=q: _dev(sd[q + "attn.to_out.0.weight"], dv, f16))
            L.b_out = take(f"l{i}.b_out", lambda q=q: _dev(sd[q + "attn.to_out.0.bias"], dv, f32))
            L.w_ff1 = take(f"l{i}.w_ff1", lambda q=q: _dev(sd[q + "ff.ff.0.0.weight"], dv, f16))
            L.b_ff1 = take(f"l{i}.b_ff1", lambda q=q: _dev(sd[q + "ff.ff.0.0.bias"], dv, f32))
            L.w_ff2 = take(f"l{i}.w_ff2", lambda q=q: _dev(sd[q + "ff.ff.2.weight"], dv, f16))
            L.b_ff2 = take(f"l{i}.b_ff2", lambda q=q: _dev(sd[q + "ff.ff.2.bias"], dv, f32))
        self._layers = layers
        w.layers = layers
        self._weights = w

        cfg = nv.DitConfig(dim, depth, heads, ff_mult, text_dim, mel_dim, self.rope_heads)
        self._cfg = cfg
        handle = nv.vp()
        nv.check(nv.load().lemas_engine_create(C.byref(cfg), C.byref(w), C.byref(handle)))
        self._handle = handle

        # RotaryEmbedding.forward_from_seq_len (x-transformers): angle[n, j] = n * inv_freq[j]; (cos, sin) pairs.
        def rope():
            inv_freq = sd.get(p + "rotary_embed.inv_freq")
            if inv_freq is None:
                inv_freq = 1.0 / (10000.0 ** (torch.arange(0, 64, 2).float() / 64))
            ang = torch.outer(torch.arange(4096, dtype=f32), inv_freq.detach().float().cpu())
            return torch.stack((ang.cos(), ang.sin()), dim=-1).to(dv).contiguous()  # [4096, 32, 2]

        take("rope", rope)
        self._rope = self._packed["rope"]
        self._ws = None
        self._traj = None  # persistent trajectory staging buffer (stable address for the step graph)

    # ------------------------------------------------------------------------------------------------ packed blob
    def export_blob(self, path) -> None:
        """Write the packed weights (what the kernels read: 0.37 GB of fp16 K-major GEMM operands + 0.55 GB of fp32
        AdaLN / time matrices for the full model) to ONE safetensors file, so that the next start uploads them as they
        are instead of re-deriving them from the 368-tensor fp32 checkpoint."""
        import json

        from safetensors.torch import save_file

        meta = dict(version=str(self.BLOB_VERSION), arch=json.dumps(dict(
            dim=self.dim, depth=self.depth, heads=self.heads, ff_mult=self.ff_mult, text_dim=self.text_dim,
            mel_dim=self.mel_dim, pe_attn_head=self.pe_attn_head)))
        save_file({k: v.detach().cpu().contiguous() for k, v in self._packed.items()}, str(path), metadata=meta)

    @classmethod
    def from_blob(cls, path, device="cuda") -> "DiTEngine":
        """Engine from a file written by `export_blob` (tensors are read straight onto the device)."""
        import json

        from safetensors import safe_open

        with safe_open(str(path), framework="pt", device=str(device)) as f:
            meta = f.metadata() or {}
            if meta.get("version") != str(cls.BLOB_VERSION):
                raise RuntimeError(f"lemas_b200 error: {path} is not a version-{cls.BLOB_VERSION} packed weight blob")
            packed = {k: f.get_tensor(k) for k in f.keys()}
        arch = json.loads(meta["arch"])
        return cls(None, device=device, prefix="", packed=packed, **arch)

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            try:
                nv.load().lemas_engine_destroy(h)
            except Exception:
                pass
            self._handle = None

    # ------------------------------------------------------------------------------------------------
    def _workspace(self, batch: int, seq: int, steps: int) -> torch.Tensor:
        need = int(nv.load().lemas_engine_workspace_bytes(C.byref(self._cfg), batch, seq, steps))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, device=self.device, dtype=torch.uint8)
        return self._ws

    def _args(self, y, step_cond, text_c, text_u, kv_len, steps, t_host, cfg_strength, traj, use_graph, flags=0,
              split=None):
        B, N, M = y.shape
        for name, t, shape in (("y", y, (B, N, self.mel_dim)), ("step_cond", step_cond, (B, N, self.mel_dim)),
                               ("text_cond", text_c, (B, N, self.text_dim))):
            if tuple(t.shape) != shape or t.dtype != f32 or not t.is_cuda or not t.is_contiguous():
                raise ValueError(f"{name}: expected contiguous CUDA fp32 {shape}, got {tuple(t.shape)} {t.dtype}")
        if text_u is not None and (tuple(text_u.shape) != (B, N, self.text_dim) or text_u.dtype != f32
                                   or not text_u.is_contiguous()):
            raise ValueError("text_uncond: bad shape/dtype")
        if kv_len is not None and (kv_len.dtype != torch.int32 or kv_len.numel() != B or not kv_len.is_cuda):
            raise ValueError("kv_len: expected CUDA int32 [batch]")
        if N > 4096:
            raise ValueError("sequence longer than 4096 frames (cfm.py:218 max_duration)")
        ws = self._workspace(B, N, max(steps, 1))
        a = nv.SampleArgs()
        a.batch, a.seq, a.steps = B, N, steps
        a.t_grid_host = t_host
        a.cfg_strength = float(cfg_strength)
        a.y, a.step_cond = nv.ptr(y), nv.ptr(step_cond)
        a.text_cond, a.text_uncond = nv.ptr(text_c), nv.ptr(text_u)
        a.kv_len = nv.ptr(kv_len)
        a.rope = nv.ptr(self._rope)
        a.trajectory = nv.ptr(traj)
        a.workspace, a.workspace_bytes = nv.ptr(ws), ws.numel()
        a.use_graph = int(use_graph)
        a.flags = int(flags)
        if split is not None:  # two-GPU CFG split (lemas_tts.parallel.CfgSplit): this process runs ONE variant
            if B * N > split.max_rows:
                raise ValueError(f"CfgSplit was set up for {split.max_rows} rows, this call has {B * N}")
            a.split_xchg_local, a.split_xchg_peer = split.xchg_ptr, split.peer_xchg_ptr
            a.split_flags_local, a.split_flags_peer = split.flags_ptr, split.peer_flags_ptr
            a.split_variant = int(split.variant)
        return a

    @nv.on_device
    def sample_loop(self, y: torch.Tensor, step_cond: torch.Tensor, text_c: torch.Tensor, text_u: torch.Tensor | None,
                    t_grid: torch.Tensor, cfg_strength: float, kv_len: torch.Tensor | None = None,
                    trajectory: torch.Tensor | None = None, use_graph: bool = True,
                    skip_padded_rows: bool = False, fold_layernorm: bool = False, split=None) -> torch.Tensor:
        """The ODE loop of CFM.sample (cfm.py:382-456).  `y` [B,N,mel] fp32 is y0 on entry and is updated IN PLACE
        to the final state.  t_grid: [steps+1] fp32 (any device; read on the host, cfm.py:445-453)."""
        tg = t_grid.detach().to("cpu", f32).contiguous()
        steps = tg.numel() - 1
        t_host = (C.c_float * (steps + 1))(*tg.tolist())
        # The step graph captures the trajectory pointer: stage the states in a persistent buffer (stable address ->
        # graph cache hit on every call) and copy them out, instead of falling back to eager launches whenever the
        # reference-default call (trajectory returned, cfm.py:456) is made.
        stage = None
        if trajectory is not None and use_graph and steps >= 3:
            n = trajectory.numel()
            if self._traj is None or self._traj.numel() < n:
                self._traj = None
                self._traj = torch.empty(n, device=self.device, dtype=f32)
            stage = self._traj[:n].view(trajectory.shape)
        a = self._args(y, step_cond, text_c, text_u, kv_len, steps, t_host, cfg_strength,
                       stage if stage is not None else trajectory, use_graph,
                       (nv.SAMPLE_SKIP_PADDED_ROWS if (skip_padded_rows and kv_len is not None) else 0)
                       | (nv.SAMPLE_FOLD_LAYERNORM if fold_layernorm else 0), split=split)
        nv.check(nv.load().lemas_sampler_run(self._handle, C.byref(a), nv.stream()))
        if stage is not None:
            trajectory.copy_(stage)
        return y

    def profile(self, enable) -> None:
        """Measurement aid (lemas_engine_profile): 1 / True = CUDA events around every eager launch, 2 = events inside
        the replayed step graph (per-kernel times of the production path), 0 / False = off."""
        nv.check(nv.load().lemas_engine_profile(self._handle, int(enable)))

    @nv.on_device
    def profile_read(self) -> dict:
        """{kind: (milliseconds, launches)} accumulated since the last read; synchronises the current stream."""
        n = len(nv.PROF_KINDS)
        ms, cnt = (C.c_double * n)(), (C.c_int64 * n)()
        nv.check(nv.load().lemas_engine_profile_read(self._handle, ms, cnt, nv.stream()))
        return {k: (ms[i], cnt[i]) for i, k in enumerate(nv.PROF_KINDS)}

    @nv.on_device
    def forward_pair(self, x: torch.Tensor, step_cond: torch.Tensor, text_c: torch.Tensor, text_u: torch.Tensor,
                     t: float, kv_len: torch.Tensor | None = None, want_hidden: bool = False):
        """One DiT.forward (dit.py:194-254) for the conditional and the unconditional variant at once.
        Returns (pred_cond, pred_uncond[, hidden]) with pred [B,N,mel] fp32."""
        B, N, M = x.shape
        t_host = (C.c_float * 2)(float(t), float(t))
        a = self._args(x, step_cond, text_c, text_u, kv_len, 1, t_host, 1.0, None, False)
        pred = torch.empty(2, B, N, 128, device=self.device, dtype=f32)
        hidden = torch.empty(2, B, N, self.dim, device=self.device, dtype=f32) if want_hidden else None
        nv.check(nv.load().lemas_dit_forward(self._handle, C.byref(a), float(t), nv.ptr(pred), nv.ptr(hidden),
                                             nv.stream()))
        out = (pred[0, ..., :M].contiguous(), pred[1, ..., :M].contiguous())
        return out + (hidden,) if want_hidden else out


class VocosEngine:
    """Packed Vocos (charactr/vocos-mel-24khz layout) weights + `lemas_vocos_decode`."""

    def __init__(self, sd: dict, device="cuda"):
        nv.require_device()
        self.device = dv = torch.device(device)
        keep = self._keep = []

        def hold(t):
            keep.append(t)
            return t

        emb = sd["backbone.embed.weight"].float()  # [dim, in_ch, 7]
        dim, in_ch, k = emb.shape
        assert k == 7 and in_ch <= 128
        n_layers = 0
        while f"backbone.convnext.{n_layers}.dwconv.weight" in sd:
            n_layers += 1
        inter = sd["backbone.convnext.0.pwconv1.weight"].shape[0]
        head = sd["head.out.weight"].float()
        if head.shape[0] != 1026:
            raise RuntimeError("lemas_b200 error: the sm_100a ISTFT head is built for n_fft=1024 (head.out rows 1026)")
        win = sd.get("head.istft.window")
        if win is not None and not torch.allclose(win.float().cpu(), torch.hann_window(1024), atol=1e-6):
            raise RuntimeError("lemas_b200 error: head.istft.window is not the periodic hann window")
        self.dim, self.inter, self.layers, self.in_ch = dim, inter, n_layers, in_ch
        w = nv.VocosWeights()
        w.dim, w.inter, w.layers, w.in_ch = dim, inter, n_layers, in_ch
        e = torch.zeros(7, dim, 128)
        e[..., :in_ch] = emb.permute(2, 0, 1)
        w.embed_w = nv.ptr(hold(_dev(e.reshape(7 * dim, 128), dv, f16)))
        w.embed_b = nv.ptr(hold(_dev(sd["backbone.embed.bias"], dv, f32)))
        w.norm_w = nv.ptr(hold(_dev(sd["backbone.norm.weight"], dv, f32)))
        w.norm_b = nv.ptr(hold(_dev(sd["backbone.norm.bias"], dv, f32)))
        blocks = (nv.VocosLayer * n_layers)()
        for i in range(n_layers):
            q = f"backbone.convnext.{i}."
            L = blocks[i]
            L.dw_w = nv.ptr(hold(_dev(sd[q + "dwconv.weight"].float()[:, 0].t(), dv, f32)))  # [7, dim]
            L.dw_b = nv.ptr(hold(_dev(sd[q + "dwconv.bias"], dv, f32)))
            L.ln_w = nv.ptr(hold(_dev(sd[q + "norm.weight"], dv, f32)))
            L.ln_b = nv.ptr(hold(_dev(sd[q + "norm.bias"], dv, f32)))
            L.w1 = nv.ptr(hold(_dev(sd[q + "pwconv1.weight"], dv, f16)))
            L.b1 = nv.ptr(hold(_dev(sd[q + "pwconv1.bias"], dv, f32)))
            L.w2 = nv.ptr(hold(_dev(sd[q + "pwconv2.weight"], dv, f16)))
            L.b2 = nv.ptr(hold(_dev(sd[q + "pwconv2.bias"], dv, f32)))
            L.gamma = nv.ptr(hold(_dev(sd[q + "gamma"], dv, f32)))
        self._blocks = blocks
        w.blocks = blocks
        w.final_w = nv.ptr(hold(_dev(sd["backbone.final_layer_norm.weight"], dv, f32)))
        w.final_b = nv.ptr(hold(_dev(sd["backbone.final_layer_norm.bias"], dv, f32)))
        hw = torch.zeros(1152, dim)
        hw[:1026] = head
        w.head_w = nv.ptr(hold(_dev(hw, dv, f16)))
        w.head_b = nv.ptr(hold(_dev(sd["head.out.bias"], dv, f32)))
        self._weights = w
        self._ws = None

    @nv.on_device
    def decode(self, mel: torch.Tensor) -> torch.Tensor:
        """Vocos.decode (utils_infer.py:549): mel [B, in_ch, T] -> wav [B, (T-1)*256] fp32."""
        if mel.dim() != 3 or mel.shape[1] != self.in_ch:
            raise ValueError(f"mel: expected [B, {self.in_ch}, T], got {tuple(mel.shape)}")
        mel = mel.to(device=self.device, dtype=f32).contiguous()
        B, _, T = mel.shape
        if T < 2:
            raise ValueError("vocos decode needs at least 2 frames")
        need = int(nv.load().lemas_vocos_workspace_bytes(C.byref(self._weights), B, T))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, device=self.device, dtype=torch.uint8)
        wav = torch.empty(B, (T - 1) * 256, device=self.device, dtype=f32)
        nv.check(nv.load().lemas_vocos_decode(C.byref(self._weights), nv.ptr(mel), nv.ptr(wav), B, T,
                                              nv.ptr(self._ws), self._ws.numel(), nv.stream()))
        return wav


class TextEngine:
    """Packed TextEmbedding weights (dit.py:34-81) + `lemas_text_embedding`."""

    def __init__(self, sd: dict, *, text_dim: int, conv_layers: int, mask_padding: bool, device="cuda",
                 prefix: str = "text_embed."):
        nv.require_device()
        self.device = dv = torch.device(device)
        keep = self._keep = []

        def hold(t):
            keep.append(t)
            return t

        p = prefix
        self.dim, self.layers = text_dim, conv_layers
        w = nv.TextWeights()
        w.dim, w.inter, w.layers, w.mask_padding = text_dim, 2 * text_dim, conv_layers, int(mask_padding)
        w.table = nv.ptr(hold(_dev(sd[p + "text_embed.weight"], dv, f32)))
        self.table_rows = int(sd[p + "text_embed.weight"].shape[0])
        if conv_layers > 0:
            inv = 1.0 / (10000.0 ** (torch.arange(0, text_dim, 2)[: text_dim // 2].float() / text_dim))
            ang = torch.outer(torch.arange(4096), inv).float()
            w.pos = nv.ptr(hold(_dev(torch.cat([ang.cos(), ang.sin()], dim=-1), dv, f32)))  # modules.py:196-207
        blocks = (nv.TextBlock * max(conv_layers, 1))()
        for i in range(conv_layers):
            q = f"{p}text_blocks.{i}."
            L = blocks[i]
            w.inter = sd[q + "pwconv1.weight"].shape[0]
            L.dw_w = nv.ptr(hold(_dev(sd[q + "dwconv.weight"].float()[:, 0].t(), dv, f32)))  # [7, dim]
            L.dw_b = nv.ptr(hold(_dev(sd[q + "dwconv.bias"], dv, f32)))
            L.ln_w = nv.ptr(hold(_dev(sd[q + "norm.weight"], dv, f32)))
            L.ln_b = nv.ptr(hold(_dev(sd[q + "norm.bias"], dv, f32)))
            L.w1 = nv.ptr(hold(_dev(sd[q + "pwconv1.weight"], dv, f16)))
            L.b1 = nv.ptr(hold(_dev(sd[q + "pwconv1.bias"], dv, f32)))
            L.grn_gamma = nv.ptr(hold(_dev(sd[q + "grn.gamma"].reshape(-1), dv, f32)))
            L.grn_beta = nv.ptr(hold(_dev(sd[q + "grn.beta"].reshape(-1), dv, f32)))
            L.w2 = nv.ptr(hold(_dev(sd[q + "pwconv2.weight"], dv, f16)))
            L.b2 = nv.ptr(hold(_dev(sd[q + "pwconv2.bias"], dv, f32)))
        self._blocks = blocks
        w.blocks = blocks
        self._weights = w
        self._ws = None

    @nv.on_device
    def embed(self, ids: torch.Tensor, drop: torch.Tensor) -> torch.Tensor:
        """ids: int32 [B, N] (shifted by +1, 0 = filler); drop: uint8 [B] -> fp32 [B, N, dim]."""
        B, N = ids.shape
        ids = ids.to(device=self.device, dtype=torch.int32).contiguous()
        drop = drop.to(device=self.device, dtype=torch.uint8).contiguous()
        # nn.Embedding in the reference (dit.py:62) faults on an id outside the table (vocab.txt that does not match the
        # checkpoint, caller-supplied token tensor); the gather kernel must never read out of bounds silently.  Same
        # contract as torch on CUDA: an asynchronous device-side assertion, no host synchronisation.
        torch._assert_async(((ids >= 0) & (ids < self.table_rows)).all(),
                            f"index out of range in TextEmbedding: ids must lie in [0, {self.table_rows})")
        need = int(nv.load().lemas_text_workspace_bytes(C.byref(self._weights), B, N))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, device=self.device, dtype=torch.uint8)
        out = torch.empty(B, N, self.dim, device=self.device, dtype=f32)
        nv.check(nv.load().lemas_text_embedding(C.byref(self._weights), nv.ptr(ids), nv.ptr(drop), nv.ptr(out), B, N,
                                                nv.ptr(self._ws), self._ws.numel(), nv.stream()))
        return out


def _pad64(c: int) -> int:
    return (c + 63) // 64 * 64


def bigvgan_aa_filter() -> torch.Tensor:
    """Kaiser-windowed sinc low-pass of BigVGAN's anti-aliased activations (alias_free_activation/torch/filter.py,
    kaiser_sinc_filter1d(cutoff=0.25, half_width=0.3, kernel_size=12)); fp32, unit DC gain."""
    import math

    k, cutoff, half_width = 12, 0.25, 0.3
    half = k // 2
    A = 2.285 * (half - 1) * math.pi * (4 * half_width) + 7.95
    beta = 0.1102 * (A - 8.7) if A > 50.0 else (0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21.0) if A >= 21.0 else 0.0)
    window = torch.kaiser_window(k, beta=beta, periodic=False)
    time = torch.arange(-half, half) + 0.5
    f = 2 * cutoff * window * torch.sinc(2 * cutoff * time)
    return (f / f.sum()).float()


def bigvgan_pack_conv(wt: torch.Tensor, cin_pad: int, cout_pad: int) -> torch.Tensor:
    """Conv1d weight [cout, cin, k] -> tap-major GEMM operand [k * cout_pad, cin_pad] (row = tap * cout_pad + out)."""
    cout, cin, k = wt.shape
    p = torch.zeros(k, cout_pad, cin_pad)
    p[:, :cout, :cin] = wt.float().permute(2, 0, 1)
    return p.reshape(k * cout_pad, cin_pad)


def bigvgan_pack_upsample(up: torch.Tensor, r: int, cin_pad: int, cout_pad: int) -> torch.Tensor:
    """ConvTranspose1d weight [cin, cout, 2r] (stride r, padding r/2) -> 3-tap GEMM operand [3 * r * cout_pad, cin_pad].
    Output sample m*r + ph takes input row m + delta (tap 0, 1, 2 = delta -1, 0, +1) through kernel index
    ph + r/2 - delta*r when that lies in [0, 2r); the other (tap, phase) blocks stay zero."""
    cin, cout, k = up.shape
    assert k == 2 * r and r % 2 == 0
    p = torch.zeros(3, r, cout_pad, cin_pad)
    for tap, delta in enumerate((-1, 0, 1)):
        for ph in range(r):
            kk = ph + r // 2 - delta * r
            if 0 <= kk < 2 * r:
                p[tap, ph, :cout, :cin] = up[:, :, kk].float().t()
    return p.reshape(3 * r * cout_pad, cin_pad)


class BigVGANEngine:
    """Packed BigVGAN-v2 weights (state dict after remove_weight_norm, NVIDIA/BigVGAN key layout) +
    `lemas_bigvgan_decode`.  Convolutions become tap-major fp16 GEMM operands on channel counts padded to 64;
    transposed convolutions become 3-tap GEMMs over (phase, channel) columns (csrc/bigvgan.cu)."""

    def __init__(self, sd: dict, h: dict, device="cuda"):
        nv.require_device()
        self.device = dv = torch.device(device)
        keep = self._keep = []

        def hold(t):
            keep.append(t)
            return t

        rates, up_k = list(h["upsample_rates"]), list(h["upsample_kernel_sizes"])
        rb_k, rb_d = list(h["resblock_kernel_sizes"]), [list(d) for d in h["resblock_dilation_sizes"]]
        if len(rb_k) != 3 or any(len(d) != 3 for d in rb_d) or str(h.get("resblock", "1")) != "1":
            raise ValueError("lemas_b200: the native BigVGAN is built for resblock '1' with 3 kernel sizes x 3 dilations")
        if h.get("activation", "snakebeta") != "snakebeta":
            raise ValueError("lemas_b200: only the 'snakebeta' activation of bigvgan_v2 is built")
        if any(k != 2 * r or r % 2 for r, k in zip(rates, up_k)):
            raise ValueError("lemas_b200: up-sampling layers must have kernel = 2 * rate and an even rate")
        logscale = bool(h.get("snake_logscale", True))
        self.num_mels = int(h["num_mels"])
        self.total_up = 1
        for r in rates:
            self.total_up *= r
        ch0 = int(h["upsample_initial_channel"])
        if ch0 % 64:
            raise ValueError("lemas_b200: upsample_initial_channel must be a multiple of 64")

        def conv_pack(wt, cin_pad, cout_pad):
            return nv.ptr(hold(_dev(bigvgan_pack_conv(wt, cin_pad, cout_pad), dv, f16)))

        def vec_pack(v, n_pad):
            p = torch.zeros(n_pad)
            p[: v.numel()] = v.float()
            return nv.ptr(hold(_dev(p, dv, f32)))

        def act_pack(prefix, c_pad):                   # [2, c_pad]: e^alpha | 1 / (e^beta + 1e-9); padded channels 0
            a, b = sd[prefix + "alpha"].float(), sd[prefix + "beta"].float()
            if logscale:
                a, b = a.exp(), b.exp()
            p = torch.zeros(2, c_pad)
            p[0, : a.numel()] = a
            p[1, : b.numel()] = 1.0 / (b + 1e-9)
            return nv.ptr(hold(_dev(p, dv, f32)))

        w = nv.BigvganWeights()
        w.num_mels, w.ch0, w.stages = self.num_mels, ch0, len(rates)
        w.use_tanh = 1 if h.get("use_tanh_at_final", True) else 0
        pre = sd["conv_pre.weight"]
        if tuple(pre.shape) != (ch0, self.num_mels, 7) or self.num_mels > 128:
            raise ValueError(f"lemas_b200: conv_pre.weight {tuple(pre.shape)} does not match the config")
        w.pre_w = conv_pack(pre, 128, ch0)
        w.pre_b = vec_pack(sd["conv_pre.bias"], ch0)
        stages = (nv.BigvganStage * len(rates))()
        cin, cin_pad = ch0, ch0
        for i, r in enumerate(rates):
            cout = cin // 2
            cpad = _pad64(cout)
            S = stages[i]
            S.rate, S.ch_in, S.ch_out = r, cin_pad, cpad
            up = sd[f"ups.{i}.0.weight"].float()        # ConvTranspose1d weight [cin, cout, 2r]
            if tuple(up.shape) != (cin, cout, 2 * r):
                raise ValueError(f"lemas_b200: ups.{i}.0.weight {tuple(up.shape)} does not match the config")
            S.up_w = nv.ptr(hold(_dev(bigvgan_pack_upsample(up, r, cin_pad, cpad), dv, f16)))
            ub = torch.zeros(r, cpad)
            ub[:, :cout] = sd[f"ups.{i}.0.bias"].float()[None, :]
            S.up_b = nv.ptr(hold(_dev(ub.reshape(-1), dv, f32)))
            for j in range(3):
                q = f"resblocks.{i * 3 + j}."
                K = S.block[j]
                K.kernel = int(rb_k[j])
                for d in range(3):
                    K.dilation[d] = int(rb_d[j][d])
                    K.w1[d] = conv_pack(sd[f"{q}convs1.{d}.weight"], cpad, cpad)
                    K.b1[d] = vec_pack(sd[f"{q}convs1.{d}.bias"], cpad)
                    K.w2[d] = conv_pack(sd[f"{q}convs2.{d}.weight"], cpad, cpad)
                    K.b2[d] = vec_pack(sd[f"{q}convs2.{d}.bias"], cpad)
                for a in range(6):
                    K.act[a] = act_pack(f"{q}activations.{a}.act.", cpad)
            cin, cin_pad = cout, cpad
        self._stages = stages
        w.stage = stages
        w.post_act = act_pack("activation_post.act.", cin_pad)
        post = sd["conv_post.weight"].float()           # [1, cin, 7]
        if tuple(post.shape) != (1, cin, 7):
            raise ValueError(f"lemas_b200: conv_post.weight {tuple(post.shape)} does not match the config")
        pw = torch.zeros(7, cin_pad)
        pw[:, :cin] = post[0].t()
        w.post_w = nv.ptr(hold(_dev(pw, dv, f32)))
        pb = sd.get("conv_post.bias")
        w.post_bias = float(pb.float().item()) if pb is not None else 0.0
        for i, v in enumerate(bigvgan_aa_filter().tolist()):
            w.aa_filter[i] = v
        self._weights = w
        self._ws = None

    @nv.on_device
    def decode(self, mel: torch.Tensor) -> torch.Tensor:
        """BigVGAN.forward (utils_infer.py:550-551): mel [B, num_mels, T] -> wav [B, 1, T * prod(rates)] fp32."""
        if mel.dim() != 3 or mel.shape[1] != self.num_mels:
            raise ValueError(f"mel: expected [B, {self.num_mels}, T], got {tuple(mel.shape)}")
        mel = mel.to(device=self.device, dtype=f32).contiguous()
        B, _, T = mel.shape
        if T < 1:
            raise ValueError("bigvgan needs at least 1 frame")
        need = int(nv.load().lemas_bigvgan_workspace_bytes(C.byref(self._weights), B, T))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, device=self.device, dtype=torch.uint8)
        wav = torch.empty(B, 1, T * self.total_up, device=self.device, dtype=f32)
        nv.check(nv.load().lemas_bigvgan_decode(C.byref(self._weights), nv.ptr(mel), nv.ptr(wav), B, T,
                                                nv.ptr(self._ws), self._ws.numel(), nv.stream()))
        return wav
