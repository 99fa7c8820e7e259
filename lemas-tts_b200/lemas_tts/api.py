"""`lemas_tts.api.TTS` — the façade the reference's entry points construct (reference lemas_tts/api.py:82-276;
callers: scripts/tts_multilingual.py:156,342, scripts/inference_gradio.py:258,297,
scripts/speech_edit_multilingual.py:388).  Same constructor / `infer` keywords, attributes and return values; the
model objects it owns (`.ema_model`, `.vocoder`) run in liblemas_b200.so.

The text/phone frontend (`lemas_tts.infer.frontend.TextNorm`: espeak-ng, jieba, pypinyin …) stays the reference's
own Python (BASELINE.json north_star); it is imported if present next to this package, see INTEGRATION.md.
"""
from __future__ import annotations

import os
import random
import sys
from pathlib import Path

from lemas_tts.infer.utils_infer import (infer_process, load_model, load_vocoder, preprocess_ref_audio_text,
                                         remove_silence_for_generated_wav, save_spectrogram, transcribe)
from lemas_tts.model.backbones.dit import DiT
from lemas_tts.model.utils import seed_everything

THIS_FILE = Path(__file__).resolve()


def _find_repo_root(start: Path) -> Path:
    for p in [start, *start.parents]:
        if (p / "pretrained_models").is_dir():
            return p
    cwd = Path.cwd()
    return cwd if (cwd / "pretrained_models").is_dir() else start


def _find_pretrained_root(start: Path) -> Path:
    """api.py:39-75: LEMAS_PRETRAINED_ROOT, then a /models/<id>/pretrained_models mount, then the source tree."""
    env_root = os.environ.get("LEMAS_PRETRAINED_ROOT")
    if env_root and Path(env_root).is_dir():
        return Path(env_root)
    models_dir = Path("/models")
    if models_dir.is_dir():
        specific = models_dir / "LEMAS-Project__LEMAS-TTS"
        if (specific / "pretrained_models").is_dir():
            return specific / "pretrained_models"
        for child in models_dir.iterdir():
            if child.is_dir() and (child / "pretrained_models").is_dir():
                return child / "pretrained_models"
    return _find_repo_root(start) / "pretrained_models"


REPO_ROOT = _find_repo_root(THIS_FILE)
PRETRAINED_ROOT = _find_pretrained_root(THIS_FILE)
CKPTS_ROOT = PRETRAINED_ROOT / "ckpts"

LANGS = {"cmn": "zh", "zh": "zh", "en": "en-us", "it": "it", "es": "es", "pt": "pt-br", "fr": "fr-fr", "de": "de",
         "ru": "ru", "id": "id", "vi": "vi", "th": "th"}
_PUNCS = {"#1", "#2", "#3", "#4", "_", "!", ",", ".", "?", '"', "'", "^", "。", "，", "？", "！"}


def load_model_config(path):
    """configs/<model>.yaml -> (arch dict, mel_spec dict).  OmegaConf when installed (api.py:99-105), else PyYAML."""
    try:
        from omegaconf import OmegaConf

        cfg = OmegaConf.to_container(OmegaConf.load(path), resolve=False)
    except ImportError:
        import yaml

        with open(path, "r") as f:
            cfg = yaml.safe_load(f)
    return dict(cfg["model"]["arch"]), dict(cfg["model"]["mel_spec"])


def process_phone_list(parts, langs=LANGS):
    """api.py:252-276: prefix every phone with the current '(lang)' tag, collapse pause marks around punctuation."""
    out, lang = [], ""
    for part in parts:
        if part.startswith("(") and part.endswith(")") and part[1:-1] in langs:
            lang = part
        elif part in _PUNCS:
            if out and out[-1] == "_":
                out.pop()
            elif out and out[-1] in _PUNCS and part == "_":
                continue
            out.append(part)
        elif lang is not None:
            out.append(f"{lang}{part}")
    return out


class TTS:
    def __init__(self, model="multilingual", ckpt_file="", vocab_file="", ode_method="euler", use_ema=False,
                 vocoder_local_path=str(CKPTS_ROOT / "vocos-mel-24khz"), use_prosody_encoder=False,
                 prosody_cfg_path="", prosody_ckpt_path="", device=None, hf_cache_dir=None, frontend="phone"):
        # `model` names a bundled config (api.py:99); a path to a yaml of the same layout is accepted as well
        cfg_file = Path(model) if str(model).endswith((".yaml", ".yml")) else THIS_FILE.parent / "configs" / f"{model}.yaml"
        model_arc, mel_cfg = load_model_config(cfg_file)
        self.mel_spec_type = mel_cfg["mel_spec_type"]
        self.target_sample_rate = mel_cfg["target_sample_rate"]
        self.ode_method = ode_method
        self.use_ema = use_ema
        self.langs = dict(LANGS)
        if device is not None:
            self.device = device
        else:
            import torch

            self.device = "cuda" if torch.cuda.is_available() else "cpu"

        vocoder_is_local = False
        if vocoder_local_path is not None:
            try:
                vocoder_is_local = Path(vocoder_local_path).is_dir()
            except TypeError:
                vocoder_is_local = False
        self.vocoder = load_vocoder(self.mel_spec_type, vocoder_is_local, vocoder_local_path, self.device, hf_cache_dir)

        if frontend is not None:
            try:
                from lemas_tts.infer.frontend import TextNorm
            except ImportError as e:
                raise ImportError("lemas_tts.infer.frontend (the reference's espeak/jieba text frontend, unchanged "
                                  "Python) is not installed next to this package — copy it from the reference tree "
                                  "or pass frontend=None and phone lists (INTEGRATION.md)") from e
            self.frontend = TextNorm(dtype=frontend)
        else:
            self.frontend = None

        self.ema_model = load_model(DiT, model_arc, ckpt_file, self.mel_spec_type, vocab_file, self.ode_method,
                                    self.use_ema, self.device, use_prosody_encoder=use_prosody_encoder,
                                    prosody_cfg_path=prosody_cfg_path, prosody_ckpt_path=prosody_ckpt_path)

    def transcribe(self, ref_audio, language=None):
        return transcribe(ref_audio, language)

    def export_wav(self, wav, file_wave, remove_silence=False):
        try:
            import soundfile as sf

            sf.write(file_wave, wav, self.target_sample_rate)
        except ImportError:
            from lemas_tts.infer.utils_infer import save_audio

            save_audio(file_wave, wav, self.target_sample_rate)
        if remove_silence:
            remove_silence_for_generated_wav(file_wave)

    def export_spectrogram(self, spec, file_spec):
        save_spectrogram(spec, file_spec)

    def infer(self, ref_file, ref_text, gen_text, show_info=print, progress=None, target_rms=0.1,
              cross_fade_duration=0.15, use_acc_grl=False, ref_ratio=None, no_ref_audio=False, cfg_strength=2,
              nfe_step=32, speed=1.0, sway_sampling_coef=5, separate_langs=False, fix_duration=None,
              use_prosody_encoder=True, file_wave=None, file_spec=None, seed=None):
        """api.py:171-249.  `ref_text` / `gen_text` are strings when a frontend is attached; with frontend=None they
        are already phone lists (`ref_text: list[str]`, `gen_text: list[list[str]]`)."""
        if progress is None:
            from lemas_tts.infer import utils_infer

            progress = utils_infer.tqdm
        if seed is None:
            seed = random.randint(0, sys.maxsize)
        seed_everything(seed)
        self.seed = seed

        if self.frontend is not None:
            ref_file, ref_text = preprocess_ref_audio_text(ref_file, ref_text)
            print("preprocesss:\n", "ref_file:", ref_file, "\nref_text:", ref_text)
            if self.frontend.dtype == "phone":
                ref_text = self.frontend.text2phn(ref_text + ". ").replace("(cmn)", "(zh)").split("|")
                gen_text = [self.frontend.text2phn(x + ". ").replace("(cmn)", "(zh)").split("|")
                            for x in gen_text.split("\n")]
            elif self.frontend.dtype == "char":
                src_lang, ref_text = self.frontend.text2norm(ref_text + ". ")
                ref_text = ["(" + src_lang.replace("cmn", "zh") + ")"] + list(ref_text)
                gen_text = [self.frontend.text2norm(x + ". ") for x in gen_text.split("\n")]
                gen_text = [["(" + x[0].replace("cmn", "zh") + ")"] + list(x[1]) for x in gen_text]
            print("after frontend:\n", "ref_text:", ref_text, "\ngen_text:", gen_text)
        if separate_langs:
            ref_text = self.process_phone_list(ref_text)
            gen_text = [self.process_phone_list(x) for x in gen_text]

        wav, sr, spec = infer_process(ref_file, ref_text, gen_text, self.ema_model, self.vocoder, self.mel_spec_type,
                                      show_info=show_info, progress=progress, target_rms=target_rms,
                                      cross_fade_duration=cross_fade_duration, nfe_step=nfe_step,
                                      cfg_strength=cfg_strength, sway_sampling_coef=sway_sampling_coef,
                                      use_prosody_encoder=use_prosody_encoder, use_acc_grl=use_acc_grl,
                                      ref_ratio=ref_ratio, no_ref_audio=no_ref_audio, speed=speed,
                                      fix_duration=fix_duration, device=self.device)
        if file_wave is not None:
            self.export_wav(wav, file_wave, remove_silence=False)
        if file_spec is not None:
            self.export_spectrogram(spec, file_spec)
        return wav, sr, spec

    def process_phone_list(self, parts):
        return process_phone_list(parts, self.langs)
