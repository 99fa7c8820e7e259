"""Short driver for ncu captures: one CFM.sample of the bench workload with a reduced step count + Vocos.decode.
usage: python tools/profile_step.py [workload=C2] [euler_steps=2]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "lemas-tts_b200")]
import torch

from lemas_tts import synthetic as syn
from lemas_tts.model.backbones.dit import DiT
from lemas_tts.model.cfm import CFM
from lemas_tts.vocoder import Vocos

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = syn.CONFIGS[name]
batch = 32 if name == "C4" else cfg.batch
arch = syn.FULL_ARCH
model = CFM(transformer=DiT(**arch.to_kwargs()), mel_spec_kwargs=dict(mel_spec_type="vocos"))
model.load_state_dict(syn.make_dit_state_dict(arch, seed=0), strict=True)
model = model.cuda()
voc = Vocos()
voc.load_state_dict(syn.make_vocos_state_dict(), strict=True)
voc = voc.cuda()
cond = syn.synthetic_ref_mel(batch, cfg.ref_frames, 100, seed=cfg.seed).cuda()
text = syn.synthetic_text_ids(batch, cfg.n_text, 898, seed=cfg.seed).cuda()
for it in range(2):
    out, _ = model.sample(cond=cond, text=text, duration=cfg.total_frames, steps=steps, cfg_strength=cfg.cfg_strength,
                          sway_sampling_coef=cfg.sway_coef, seed=1, use_acc_grl=False, return_trajectory=False)
    wav = voc.decode(out[:, cfg.ref_frames:, :].permute(0, 2, 1))
torch.cuda.synchronize()
print("done", tuple(out.shape), tuple(wav.shape))
