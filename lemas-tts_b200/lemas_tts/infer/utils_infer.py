"""Inference glue behind the reference's `lemas_tts.infer.utils_infer` surface (reference
lemas_tts/infer/utils_infer.py): module constants, load_vocoder / load_checkpoint / load_model,
preprocess_ref_audio_text, infer_process and infer_batch_process with the same signatures, defaults and yields.

The model objects these return run their hot path in liblemas_b200.so (CFM.sample -> lemas_sampler_run,
vocoder.decode -> lemas_vocos_decode).  Host-only helpers of the reference that need packages absent offline
(pydub silence trimming, Whisper ASR, matplotlib) import lazily and raise ImportError with the missing name.
"""
from __future__ import annotations

import hashlib
import os
import re
from pathlib import Path

import numpy as np
import torch
import torchaudio

from lemas_tts.model.cfm import CFM
from lemas_tts.model.utils import convert_char_to_pinyin, get_tokenizer
from lemas_tts.vocoder import Vocos

try:  # tqdm is what the reference passes as `progress`
    import tqdm
except ImportError:  # pragma: no cover
    tqdm = None


def load_audio(path):
    """torchaudio.load, falling back to scipy's WAV reader when torchaudio has no decoding backend installed
    (torchaudio >= 2.9 delegates to torchcodec).  Returns (float32 [channels, samples], sample_rate)."""
    try:
        return torchaudio.load(path)
    except (ImportError, RuntimeError, OSError):
        from scipy.io import wavfile

        sr, data = wavfile.read(path)
        x = torch.from_numpy(np.ascontiguousarray(data))
        if x.dtype == torch.int16:
            x = x.float() / 32768.0
        elif x.dtype == torch.int32:
            x = x.float() / 2147483648.0
        elif x.dtype == torch.uint8:
            x = (x.float() - 128.0) / 128.0
        else:
            x = x.float()
        x = x[None] if x.dim() == 1 else x.t().contiguous()
        return x, int(sr)


def save_audio(path, wave, sample_rate):
    """Write float32 [channels, samples] (or [samples]) as 16-bit PCM WAV; torchaudio when it can, else scipy."""
    wave = torch.as_tensor(wave, dtype=torch.float32)
    wave = wave[None] if wave.dim() == 1 else wave
    try:
        torchaudio.save(path, wave, sample_rate)
    except (ImportError, RuntimeError, OSError):
        from scipy.io import wavfile

        pcm = (wave.clamp(-1, 1) * 32767.0).round().to(torch.int16).t().contiguous().numpy()
        wavfile.write(path, int(sample_rate), pcm[:, 0] if pcm.shape[1] == 1 else pcm)


def _find_repo_root(start: Path) -> Path:
    for p in [start, *start.parents]:
        if (p / "pretrained_models").is_dir():
            return p
    cwd = Path.cwd()
    if (cwd / "pretrained_models").is_dir():
        return cwd
    return start


THIS_FILE = Path(__file__).resolve()
REPO_ROOT = _find_repo_root(THIS_FILE)
PRETRAINED_ROOT = REPO_ROOT / "pretrained_models"
CKPTS_ROOT = PRETRAINED_ROOT / "ckpts"

_ref_audio_cache: dict = {}

device = "cuda" if torch.cuda.is_available() else "cpu"

# ----------------------------------------- utils_infer.py:68-81
target_sample_rate = 24000
n_mel_channels = 100
hop_length = 256
win_length = 1024
n_fft = 1024
mel_spec_type = "vocos"
target_rms = 0.1
cross_fade_duration = 0.15
ode_method = "euler"
nfe_step = 32
cfg_strength = 3.0
sway_sampling_coef = 1
speed = 1.0
fix_duration = None
# -----------------------------------------


def chunk_text(text, max_chars=135):
    """utils_infer.py:89-116: greedy sentence packing under a UTF-8 byte budget."""
    chunks, current = [], ""
    for sentence in re.split(r"(?<=[;:,.!?])\s+|(?<=[；：，。！？])", text):
        piece = sentence + " " if sentence and len(sentence[-1].encode("utf-8")) == 1 else sentence
        if len(current.encode("utf-8")) + len(sentence.encode("utf-8")) <= max_chars:
            current += piece
        else:
            if current:
                chunks.append(current.strip())
            current = piece
    if current:
        chunks.append(current.strip())
    return chunks


def load_vocoder(vocoder_name="vocos", is_local=False, local_path="", device=device, hf_cache_dir=None):
    """utils_infer.py:120-159.  vocos: an object with `.decode(mel[B,100,T]) -> wav[B,S]`; bigvgan: a callable
    `vocoder(mel[B,100,T]) -> wav[B,1,S]`."""
    if vocoder_name == "bigvgan":
        # utils_infer.py:144-158: BigVGAN.from_pretrained(local dir | hub id), remove_weight_norm, eval, to(device);
        # the un-vendored third_party/BigVGAN class is replaced by lemas_tts.bigvgan.BigVGAN (same surface)
        from lemas_tts.bigvgan import BigVGAN

        if is_local:
            vocoder = BigVGAN.from_pretrained(local_path, use_cuda_kernel=False)
        else:
            vocoder = BigVGAN.from_pretrained("nvidia/bigvgan_v2_24khz_100band_256x", use_cuda_kernel=False,
                                              cache_dir=hf_cache_dir)
        vocoder.remove_weight_norm()
        return vocoder.eval().to(device)
    if vocoder_name != "vocos":
        raise ValueError(f"unknown vocoder {vocoder_name!r} (vocos | bigvgan)")
    if is_local:
        print(f"Load vocos from local path {local_path}")
        config_path = f"{local_path}/config.yaml"
        model_path = f"{local_path}/pytorch_model.bin"
    else:
        print("Download Vocos from huggingface charactr/vocos-mel-24khz")
        from huggingface_hub import hf_hub_download

        repo_id = "charactr/vocos-mel-24khz"
        config_path = hf_hub_download(repo_id=repo_id, cache_dir=hf_cache_dir, filename="config.yaml")
        model_path = hf_hub_download(repo_id=repo_id, cache_dir=hf_cache_dir, filename="pytorch_model.bin")
    vocoder = Vocos.from_hparams(config_path)
    state_dict = torch.load(model_path, map_location="cpu", weights_only=True)
    vocoder.load_state_dict(state_dict)
    return vocoder.eval().to(device)


asr_pipe = None


def initialize_asr_pipeline(device: str = device, dtype=None):
    """utils_infer.py:167-184 (Whisper; host-side, only when no reference text is given)."""
    from transformers import pipeline

    if dtype is None:
        dtype = torch.float16 if "cuda" in str(device) else torch.float32
    global asr_pipe
    asr_pipe = pipeline("automatic-speech-recognition", model="openai/whisper-large-v3-turbo", torch_dtype=dtype,
                        device=device)


def transcribe(ref_audio, language=None):
    global asr_pipe
    if asr_pipe is None:
        initialize_asr_pipeline(device=device)
    return asr_pipe(ref_audio, chunk_length_s=30, batch_size=128,
                    generate_kwargs={"task": "transcribe", "language": language} if language else {"task": "transcribe"},
                    return_timestamps=False)["text"].strip()


_LEGACY_KEYS = ["mel_spec.mel_stft.mel_scale.fb", "mel_spec.mel_stft.spectrogram.window", "ctc.proj.0.weight",
                "ctc.proj.0.bias", "ctc.ctc_proj.weight", "ctc.ctc_proj.bias"]


def load_checkpoint(model, ckpt_path, device: str, dtype=None, use_ema=True):
    """utils_infer.py:204-246: same file formats, key remapping and STRICT key check.

    `dtype` is accepted for signature compatibility.  The reference casts the module to fp16 on CUDA; here the
    parameter containers stay fp32 masters and the engine packs fp16 tensor-core operands from them with fp32
    accumulation and an fp32 residual stream / ODE state (>= the reference's precision).
    """
    ckpt_path = str(ckpt_path)
    ckpt_type = ckpt_path.split(".")[-1]
    if ckpt_type == "safetensors":
        from safetensors.torch import load_file

        checkpoint = load_file(ckpt_path, device="cpu")
    else:
        checkpoint = torch.load(ckpt_path, map_location="cpu", weights_only=True)

    if use_ema:
        if ckpt_type == "safetensors":
            checkpoint = {"ema_model_state_dict": checkpoint}
        state = {k.replace("ema_model.", ""): v for k, v in checkpoint["ema_model_state_dict"].items()
                 if k not in ["initted", "step"]}
        for key in _LEGACY_KEYS:
            state.pop(key, None)
    else:
        if ckpt_type == "safetensors":
            checkpoint = {"model_state_dict": checkpoint}
        state = checkpoint["model_state_dict"]
    model.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in state.items()})
    del checkpoint
    return model.to(device)


def load_model(model_cls, model_cfg, ckpt_path, mel_spec_type=mel_spec_type, vocab_file="", ode_method=ode_method,
               use_ema=True, device=device, use_prosody_encoder=False, prosody_cfg_path="", prosody_ckpt_path=""):
    """utils_infer.py:252-303."""
    if vocab_file == "":
        from importlib.resources import files

        vocab_file = str(files("lemas_tts").joinpath("infer/examples/vocab.txt"))
    tokenizer = "custom"
    print("\nvocab : ", vocab_file)
    print("token : ", tokenizer)
    print("model : ", ckpt_path, "\n")
    vocab_char_map, vocab_size = get_tokenizer(vocab_file, tokenizer)
    if use_prosody_encoder:
        if not prosody_cfg_path:
            prosody_cfg_path = str(CKPTS_ROOT / "prosody_encoder" / "pretssel_cfg.json")
        if not prosody_ckpt_path:
            prosody_ckpt_path = str(CKPTS_ROOT / "prosody_encoder" / "prosody_encoder_UnitY2.pt")
    model = CFM(
        transformer=model_cls(**model_cfg, text_num_embeds=vocab_size, mel_dim=n_mel_channels,
                              use_prosody_encoder=use_prosody_encoder),
        mel_spec_kwargs=dict(n_fft=n_fft, hop_length=hop_length, win_length=win_length, n_mel_channels=n_mel_channels,
                             target_sample_rate=target_sample_rate, mel_spec_type=mel_spec_type),
        odeint_kwargs=dict(method=ode_method),
        vocab_char_map=vocab_char_map,
        use_prosody_encoder=use_prosody_encoder,
        prosody_cfg_path=prosody_cfg_path,
        prosody_ckpt_path=prosody_ckpt_path,
    ).to(device)
    dtype = torch.float32 if mel_spec_type == "bigvgan" else None
    return load_checkpoint(model, ckpt_path, device, dtype=dtype, use_ema=use_ema)


# ----------------------------------------------------------------------------- reference-audio preprocessing


def _dbfs(x: torch.Tensor) -> float:
    rms = float(x.float().pow(2).mean().sqrt()) if x.numel() else 0.0
    return 20.0 * float(np.log10(max(rms, 1e-10)))


def remove_silence_edges(wave: torch.Tensor, sr: int, silence_threshold=-42) -> torch.Tensor:
    """utils_infer.py:306-320 on a [channels, samples] tensor: strip leading (10 ms steps) and trailing (1 ms steps)
    audio quieter than `silence_threshold` dBFS."""
    step = max(1, sr // 100)
    start = 0
    while start < wave.shape[-1] and _dbfs(wave[..., start:start + step]) < silence_threshold:
        start += step
    wave = wave[..., start:]
    ms = max(1, sr // 1000)
    end = wave.shape[-1]
    while end > 0 and _dbfs(wave[..., max(0, end - ms):end]) <= silence_threshold:
        end -= ms
    return wave[..., :max(end, 0)]


def preprocess_ref_audio_text(ref_audio_orig, ref_text, clip_short=True, show_info=print):
    """utils_infer.py:325-393: clip the reference to <= 12 s, trim edge silence, append 50 ms of silence, cache the
    transcription by audio hash, and make the reference text end in '. '.  Uses pydub exactly like the reference
    when it is installed; otherwise a torchaudio restatement of the clip / trim steps."""
    import tempfile

    show_info("Converting audio...")
    try:
        from pydub import AudioSegment, silence  # noqa: F401

        have_pydub = True
    except ImportError:
        have_pydub = False
    with tempfile.NamedTemporaryFile(delete=False, suffix=".wav") as f:
        out_path = f.name
    if have_pydub:
        aseg = AudioSegment.from_file(ref_audio_orig)
        if clip_short:
            for min_len, thresh, tag in ((1000, -50, "(1)"), (100, -40, "(2)")):
                segs = silence.split_on_silence(aseg, min_silence_len=min_len, silence_thresh=thresh,
                                                keep_silence=1000, seek_step=10)
                wave = AudioSegment.silent(duration=0)
                for seg in segs:
                    if len(wave) > 6000 and len(wave + seg) > 12000:
                        show_info(f"Audio is over 12s, clipping short. {tag}")
                        break
                    wave += seg
                if len(wave) <= 12000:
                    break
            aseg = wave
            if len(aseg) > 12000:
                aseg = aseg[:12000]
                show_info("Audio is over 12s, clipping short. (3)")
        start = silence.detect_leading_silence(aseg, silence_threshold=-42)
        aseg = aseg[start:]
        end = aseg.duration_seconds
        for ms in reversed(aseg):
            if ms.dBFS > -42:
                break
            end -= 0.001
        aseg = aseg[: int(end * 1000)] + AudioSegment.silent(duration=50)
        aseg.export(out_path, format="wav")
    else:
        wave, sr = load_audio(ref_audio_orig)
        if clip_short and wave.shape[-1] > 12 * sr:
            wave = wave[..., : 12 * sr]
            show_info("Audio is over 12s, clipping short. (3)")
        wave = remove_silence_edges(wave, sr)
        wave = torch.cat([wave, torch.zeros(wave.shape[0], int(0.05 * sr))], dim=-1)
        save_audio(out_path, wave, sr)
    ref_audio = out_path

    with open(ref_audio, "rb") as audio_file:
        audio_hash = hashlib.md5(audio_file.read()).hexdigest()
    if not ref_text.strip():
        if audio_hash in _ref_audio_cache:
            show_info("Using cached reference text...")
            ref_text = _ref_audio_cache[audio_hash]
        else:
            show_info("No reference text provided, transcribing reference audio...")
            ref_text = transcribe(ref_audio)
            _ref_audio_cache[audio_hash] = ref_text
    else:
        show_info("Using custom reference text...")
    if not ref_text.endswith(". ") and not ref_text.endswith("。"):
        ref_text += " " if ref_text.endswith(".") else ". "
    print("\nref_text  ", ref_text)
    return ref_audio, ref_text


# ----------------------------------------------------------------------------- inference


def infer_process(ref_audio, ref_text, gen_text, model_obj, vocoder, mel_spec_type=mel_spec_type, show_info=print,
                  progress=tqdm, target_rms=target_rms, cross_fade_duration=cross_fade_duration, nfe_step=nfe_step,
                  cfg_strength=cfg_strength, sway_sampling_coef=sway_sampling_coef, use_acc_grl=True,
                  use_prosody_encoder=True, ref_ratio=None, no_ref_audio=False, speed=speed, fix_duration=fix_duration,
                  device=device):
    """utils_infer.py:399-458: chunk the text (string input) and run infer_batch_process once."""
    audio, sr = load_audio(ref_audio)
    if type(ref_text) == str:
        secs = audio.shape[-1] / sr
        max_chars = int(len(ref_text.encode("utf-8")) / secs * (22 - secs))
        gen_text_batches = chunk_text(gen_text, max_chars=max_chars)
    else:
        gen_text_batches = gen_text
    print("ref_text:", ref_text)
    for i, g in enumerate(gen_text_batches):
        print(f"gen_text {i}", g)
    print("\n")
    show_info(f"Generating audio in {len(gen_text_batches)} batches...")
    return next(infer_batch_process((audio, sr), ref_text, gen_text_batches, model_obj, vocoder,
                                    mel_spec_type=mel_spec_type, progress=progress, target_rms=target_rms,
                                    cross_fade_duration=cross_fade_duration, nfe_step=nfe_step,
                                    cfg_strength=cfg_strength, sway_sampling_coef=sway_sampling_coef,
                                    use_acc_grl=use_acc_grl, use_prosody_encoder=use_prosody_encoder,
                                    ref_ratio=ref_ratio, no_ref_audio=no_ref_audio, speed=speed,
                                    fix_duration=fix_duration, device=device))


def cross_fade_concat(waves: list, cross_fade_duration: float, sample_rate: int = target_sample_rate) -> np.ndarray:
    """utils_infer.py:581-617: linear cross-fade of consecutive chunks over `cross_fade_duration` seconds."""
    if cross_fade_duration <= 0:
        return np.concatenate(waves)
    final = waves[0]
    for nxt in waves[1:]:
        n = min(int(cross_fade_duration * sample_rate), len(final), len(nxt))
        if n <= 0:
            final = np.concatenate([final, nxt])
            continue
        mixed = final[-n:] * np.linspace(1, 0, n) + nxt[:n] * np.linspace(0, 1, n)
        final = np.concatenate([final[:-n], mixed, nxt[n:]])
    return final


def infer_batch_process(ref_audio, ref_text, gen_text_batches, model_obj, vocoder, mel_spec_type="vocos",
                        progress=tqdm, target_rms=0.1, cross_fade_duration=0.15, nfe_step=32, cfg_strength=2.0,
                        sway_sampling_coef=-1, use_acc_grl=True, use_prosody_encoder=True, ref_ratio=None,
                        no_ref_audio=False, speed=1, fix_duration=None, device=None, streaming=False, chunk_size=2048):
    """utils_infer.py:464-625 (generator).  Non-streaming: yields (final_wave, 24000, combined_spectrogram) once;
    streaming: yields (chunk, 24000) pieces of `chunk_size` samples."""
    audio, sr = ref_audio
    # On a CUDA device the waveform-side arithmetic runs in csrc/audio.cu / csrc/prosody.cu (mono mix, RMS scaling,
    # sinc resampling, un-scaling, cross-fade, clip): the raw reference audio goes up once, the final waveform comes
    # down once, nothing in between synchronises with the host.
    on_device = torch.device(device if device is not None else "cpu").type == "cuda"
    stats = None
    if on_device:
        from lemas_tts import audio_native, prosody_native

        audio, stats = audio_native.prep_reference_audio(audio.to(device), target_rms)
        if sr != target_sample_rate:
            audio = prosody_native.resample(audio, sr, target_sample_rate)
    else:
        if audio.shape[0] > 1:
            audio = torch.mean(audio, dim=0, keepdim=True)
        rms = torch.sqrt(torch.mean(torch.square(audio)))
        if rms < target_rms:
            audio = audio * target_rms / rms
        if sr != target_sample_rate:
            audio = torchaudio.transforms.Resample(sr, target_sample_rate)(audio)
        audio = audio.to(device)

    if type(ref_text) == str and len(ref_text[-1].encode("utf-8")) == 1:
        ref_text = ref_text + " "

    def process_batch(gen_text):
        local_speed = speed
        if type(ref_text) == str:
            if len(gen_text.encode("utf-8")) < 10:
                local_speed = 0.3
            final_text_list = convert_char_to_pinyin([ref_text + gen_text])
        else:
            final_text_list = [ref_text + gen_text]
        print("final_text_list:", final_text_list)

        ref_audio_len = audio.shape[-1] // hop_length
        if fix_duration is not None:
            duration = int(fix_duration * target_sample_rate / hop_length)
        else:
            duration = ref_audio_len + int(ref_audio_len / len(ref_text) * len(gen_text) / local_speed)

        with torch.inference_mode():
            generated, _ = model_obj.sample(cond=audio, text=final_text_list, duration=duration, steps=nfe_step,
                                            cfg_strength=cfg_strength, sway_sampling_coef=sway_sampling_coef,
                                            use_acc_grl=use_acc_grl, use_prosody_encoder=use_prosody_encoder,
                                            ref_ratio=ref_ratio, no_ref_audio=no_ref_audio,
                                            return_trajectory=False)
            del _
            generated = generated.to(torch.float32)[:, ref_audio_len:, :].permute(0, 2, 1)
            if mel_spec_type == "vocos":
                wave = vocoder.decode(generated)
            else:
                wave = vocoder(generated)
            if on_device:
                wave = audio_native.unscale_(wave.contiguous(), stats).squeeze()
                if streaming:
                    wave = wave.cpu().numpy()
            else:
                if rms < target_rms:
                    wave = wave * rms / target_rms
                wave = wave.squeeze().cpu().numpy()
            if streaming:
                for j in range(0, len(wave), chunk_size):
                    yield wave[j:j + chunk_size], target_sample_rate
            else:
                yield wave, generated[0].cpu().numpy()

    if streaming:
        for gen_text in progress.tqdm(gen_text_batches) if progress is not None else gen_text_batches:
            for chunk in process_batch(gen_text):
                yield chunk
        return

    waves, specs = [], []
    # The reference submits process_batch (a generator function) to a ThreadPoolExecutor, so the work runs serially
    # in the caller at next(result) anyway; chunks are processed in order here.
    batches = progress.tqdm(gen_text_batches) if progress is not None else gen_text_batches
    for gen_text in batches:
        wave, spec = next(process_batch(gen_text))
        waves.append(wave)
        specs.append(spec)
    if waves:
        if on_device:
            final_wave = audio_native.cross_fade_concat(waves, cross_fade_duration, target_sample_rate, clip=0.999)
        else:
            final_wave = np.clip(cross_fade_concat(waves, cross_fade_duration), -0.999, 0.999)
        yield final_wave, target_sample_rate, np.concatenate(specs, axis=1)
    else:
        yield None, target_sample_rate, None


def remove_silence_for_generated_wav(filename):
    """utils_infer.py:631-640 (pydub)."""
    from pydub import AudioSegment, silence

    aseg = AudioSegment.from_file(filename)
    out = AudioSegment.silent(duration=0)
    for seg in silence.split_on_silence(aseg, min_silence_len=1000, silence_thresh=-50, keep_silence=500, seek_step=10):
        out += seg
    out.export(filename, format="wav")


def save_spectrogram(spectrogram, path):
    """utils_infer.py:646-651 (matplotlib)."""
    import matplotlib

    matplotlib.use("Agg")
    import matplotlib.pylab as plt

    plt.figure(figsize=(12, 4))
    plt.imshow(spectrogram, origin="lower", aspect="auto")
    plt.colorbar()
    plt.savefig(path)
    plt.close()
