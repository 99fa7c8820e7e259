"""Time the prosody path (resample -> kaldi fbank -> ECAPA-TDNN) on the GPU box: native fp32 kernels vs the same
arithmetic in torch / torchaudio ops on the device.  C3 shape: 32 utterances x 10 s of 24 kHz audio, full-size encoder."""
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "lemas-tts_b200")]
import torch
import torchaudio

from lemas_tts import prosody_native as pn
from lemas_tts import synthetic as syn
from lemas_tts.model.backbones.prosody_encoder import ProsodyEncoder

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
with tempfile.TemporaryDirectory() as tmp:
    cfg_path, ckpt_path = syn.write_prosody_assets(Path(tmp), syn.PROSODY_CFG, seed=13)
    enc = ProsodyEncoder(cfg_path, ckpt_path).eval().cuda()
raw = syn.synthetic_ref_audio(B, 240000, seed=1).cuda()


def native():
    return enc(pn.kaldi_fbank_80(pn.resample(raw, 24000, 16000)))


def torch_ops():
    a16 = torchaudio.functional.resample(raw, 24000, 16000)
    fb = torch.stack([torchaudio.compliance.kaldi.fbank(a[None], num_mel_bins=80, sample_frequency=16000) for a in a16])
    return enc.encoder.forward_torch(fb)


with torch.no_grad():
    for name, fn in (("native csrc/prosody.cu", native), ("torch/torchaudio ops on the device", torch_ops)):
        for _ in range(2):
            out = fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            out = fn()
        torch.cuda.synchronize()
        print(f"{name:40s} {1e3 * (time.perf_counter() - t0) / 5:8.2f} ms per batch of {B} x 10 s")
    print("max |native - torch|:", (native() - torch_ops()).abs().max().item())
