"""Run the memory-bound kernels of the path a few times (for `ncu --set full -k regex:...`):
LN+modulate at the C2 and C4 shapes, Vocos decode (dwconv7+LN, iSTFT), mel front-end, kaldi fbank."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "lemas-tts_b200")]
import torch
from lemas_tts import ops, prosody_native as pn, synthetic as syn
from lemas_tts.vocoder import Vocos

g = torch.Generator(device="cuda").manual_seed(0)
voc = Vocos(); voc.load_state_dict(syn.make_vocos_state_dict(), strict=True); voc = voc.cuda()
mel = torch.randn(1, 100, 1250, device="cuda", generator=g)
wav = syn.synthetic_ref_audio(1, 240000, seed=1).cuda()
for rows in (4374, 49152):
    x = torch.randn(rows, 1024, device="cuda", generator=g)
    sc = torch.randn(1024, device="cuda", generator=g) * 0.1
    for _ in range(3):
        ops.ln_modulate(x, sc, sc, rows)
for _ in range(3):
    voc.decode(mel)
    ops.mel_spectrogram_1024(wav)
    pn.kaldi_fbank_80(pn.resample(wav, 24000, 16000))
torch.cuda.synchronize()
