"""Drop-in surface end to end on the GPU: `lemas_tts.api.TTS` built from checkpoint / vocab / vocoder files in the
reference's formats (safetensors with `ema_model.` keys, vocab.txt, vocos config.yaml + pytorch_model.bin), then
`TTS.infer` -> `infer_process` -> `infer_batch_process` -> `CFM.sample` + `vocoder.decode` -> cross-faded waveform —
the call chain of scripts/tts_multilingual.py:342 (reference), with the text frontend replaced by phone lists."""
import numpy as np
import pytest
import torch

from lemas_tts import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def assets(tmp_path_factory):
    import yaml
    from safetensors.torch import save_file

    from lemas_tts.infer.utils_infer import save_audio

    d = tmp_path_factory.mktemp("pretrained")
    arch = syn.TINY_ARCH
    cfg = dict(model=dict(arch=dict(dim=arch.dim, depth=arch.depth, heads=arch.heads, ff_mult=arch.ff_mult,
                                    text_dim=arch.text_dim, text_mask_padding=True, qk_norm=None,
                                    conv_layers=arch.conv_layers, pe_attn_head=None, checkpoint_activations=False),
                          mel_spec=dict(target_sample_rate=24000, n_mel_channels=100, hop_length=256, win_length=1024,
                                        n_fft=1024, mel_spec_type="vocos")))
    (d / "tiny.yaml").write_text(yaml.safe_dump(cfg))
    vocab = [" "] + [f"p{i}" for i in range(1, arch.text_num_embeds)]
    (d / "vocab.txt").write_text("\n".join(vocab) + "\n")
    sd = syn.make_dit_state_dict(arch, seed=11)
    ema = {"ema_model." + k: v.contiguous() for k, v in sd.items()}
    ema["initted"], ema["step"] = torch.tensor(1.0), torch.tensor(1.0)
    save_file(ema, str(d / "model.safetensors"))
    voc = d / "vocos-mel-24khz"
    voc.mkdir()
    va = syn.TINY_VOCOS
    (voc / "config.yaml").write_text(yaml.safe_dump(dict(
        backbone=dict(class_path="vocos.models.VocosBackbone",
                      init_args=dict(input_channels=100, dim=va.dim, intermediate_dim=va.intermediate_dim,
                                     num_layers=va.num_layers)),
        head=dict(class_path="vocos.heads.ISTFTHead", init_args=dict(dim=va.dim, n_fft=1024, hop_length=256,
                                                                     padding="center")))))
    torch.save(syn.make_vocos_state_dict(va, seed=7), voc / "pytorch_model.bin")
    save_audio(str(d / "ref.wav"), syn.synthetic_ref_audio(1, 36000, seed=3), 24000)  # 1.5 s
    return d, vocab


def test_tts_infer_end_to_end(assets):
    from lemas_tts.api import TTS

    d, vocab = assets
    tts = TTS(model=str(d / "tiny.yaml"), ckpt_file=str(d / "model.safetensors"), vocab_file=str(d / "vocab.txt"),
              use_ema=True, vocoder_local_path=str(d / "vocos-mel-24khz"), device="cuda", frontend=None)
    assert tts.mel_spec_type == "vocos" and tts.target_sample_rate == 24000
    ref_text = vocab[1:21]
    gen_text = [vocab[5:35], vocab[10:25] + ["never-seen-token"]]  # two chunks -> cross-fade; unknown token -> id 0
    wav, sr, spec = tts.infer(str(d / "ref.wav"), ref_text, gen_text, nfe_step=4, cfg_strength=2, sway_sampling_coef=5,
                              seed=123, progress=None)
    assert sr == 24000 and isinstance(wav, np.ndarray) and wav.ndim == 1
    assert np.isfinite(wav).all() and np.abs(wav).max() <= 0.999 and np.abs(wav).max() > 0
    ref_frames = 36000 // 256
    exp_frames = [int(ref_frames / 20 * 30), int(ref_frames / 20 * 16)]  # utils_infer.py:520-527
    assert spec.shape == (100, sum(exp_frames))
    exp_len = sum((f - 1) * 256 for f in exp_frames) - int(0.15 * 24000)  # one cross-fade
    assert len(wav) == exp_len
    wav2, _, _ = tts.infer(str(d / "ref.wav"), ref_text, gen_text, nfe_step=4, cfg_strength=2, sway_sampling_coef=5,
                           seed=123, progress=None)
    assert np.array_equal(wav, wav2), "same seed -> same waveform"


def test_streaming_yields_chunks(assets):
    from lemas_tts.api import TTS
    from lemas_tts.infer.utils_infer import infer_batch_process, load_audio

    d, vocab = assets
    tts = TTS(model=str(d / "tiny.yaml"), ckpt_file=str(d / "model.safetensors"), vocab_file=str(d / "vocab.txt"),
              use_ema=True, vocoder_local_path=str(d / "vocos-mel-24khz"), device="cuda", frontend=None)
    audio, sr = load_audio(str(d / "ref.wav"))
    chunks = list(infer_batch_process((audio, sr), vocab[1:21], [vocab[5:35]], tts.ema_model, tts.vocoder,
                                      progress=None, nfe_step=2, device="cuda", streaming=True, chunk_size=2048))
    assert all(sr_ == 24000 for _, sr_ in chunks) and all(len(c) <= 2048 for c, _ in chunks)
    assert sum(len(c) for c, _ in chunks) == (int(36000 // 256 / 20 * 30) - 1) * 256


def test_tts_infer_with_the_bigvgan_branch(assets, tmp_path):
    """`mel_spec_type: bigvgan` (configs/*.yaml:65 "vocos | bigvgan"): the Slaney-mel front-end for the reference audio
    (modules.py:30-72), `load_vocoder("bigvgan", local dir)` (utils_infer.py:144-158) and `vocoder(mel)` (:550-551)."""
    import json

    import yaml

    from lemas_tts.api import TTS

    d, vocab = assets
    cfg = yaml.safe_load((d / "tiny.yaml").read_text())
    cfg["model"]["mel_spec"]["mel_spec_type"] = "bigvgan"
    (tmp_path / "tiny_bigvgan.yaml").write_text(yaml.safe_dump(cfg))
    import dataclasses

    arch = dataclasses.replace(syn.FULL_BIGVGAN, upsample_initial_channel=128)   # 256x up-sampling, 64 ... 2 channels
    voc = tmp_path / "bigvgan_v2_24khz_100band_256x"
    voc.mkdir()
    (voc / "config.json").write_text(json.dumps(arch.to_config()))
    torch.save({"generator": syn.make_bigvgan_state_dict(arch, seed=5)}, voc / "bigvgan_generator.pt")
    tts = TTS(model=str(tmp_path / "tiny_bigvgan.yaml"), ckpt_file=str(d / "model.safetensors"),
              vocab_file=str(d / "vocab.txt"), use_ema=True, vocoder_local_path=str(voc), device="cuda", frontend=None)
    assert tts.mel_spec_type == "bigvgan"
    ref_text = vocab[1:21]
    gen_text = [vocab[5:35], vocab[10:26]]
    wav, sr, spec = tts.infer(str(d / "ref.wav"), ref_text, gen_text, nfe_step=4, cfg_strength=2, sway_sampling_coef=5,
                              seed=123, progress=None)
    assert sr == 24000 and wav.ndim == 1 and np.isfinite(wav).all() and np.abs(wav).max() > 0
    ref_frames = 36000 // 256
    exp_frames = [int(ref_frames / 20 * 30), int(ref_frames / 20 * 16)]
    assert spec.shape == (100, sum(exp_frames))
    assert len(wav) == sum(f * 256 for f in exp_frames) - int(0.15 * 24000)       # BigVGAN: 256 samples per frame
