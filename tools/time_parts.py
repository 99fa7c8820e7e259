"""Where the per-utterance time goes outside the ODE loop (C2 shape): sample(steps) for several step counts, text
embedding, vocoder."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "lemas-tts_b200")]
import torch
from lemas_tts import synthetic as syn
from lemas_tts.model.backbones.dit import DiT
from lemas_tts.model.cfm import CFM
from lemas_tts.vocoder import Vocos
cfg = syn.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C2"]
arch = syn.FULL_ARCH
model = CFM(transformer=DiT(**arch.to_kwargs()), mel_spec_kwargs=dict(mel_spec_type="vocos"))
model.load_state_dict(syn.make_dit_state_dict(arch, seed=0), strict=True)
model = model.cuda()
voc = Vocos(); voc.load_state_dict(syn.make_vocos_state_dict(), strict=True); voc = voc.cuda()
cond = syn.synthetic_ref_mel(1, cfg.ref_frames, 100, seed=0).cuda()
text = syn.synthetic_text_ids(1, cfg.n_text, 898, seed=0).cuda()

def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

res = {}
for steps in (3, 8, 32):
    res[steps] = t(lambda: model.sample(cond=cond, text=text, duration=cfg.total_frames, steps=steps, cfg_strength=2.0,
                                        sway_sampling_coef=5.0, seed=1, use_acc_grl=False, return_trajectory=False))
    print(f"sample steps={steps}: {res[steps]:.2f} ms")
per_step = (res[32] - res[8]) / 24
print(f"per ODE step {per_step:.3f} ms; fixed per-call overhead {res[32] - 32 * per_step:.2f} ms")
tr = model.transformer
print(f"text embeds (2 passes, torch): {t(lambda: tr.text_embeds(text, cfg.total_frames, None, cache=False)):.2f} ms")
mel = torch.randn(1, 100, cfg.total_frames - cfg.ref_frames, device='cuda')
print(f"vocos decode {mel.shape[-1]} frames: {t(lambda: voc.decode(mel)):.2f} ms")
