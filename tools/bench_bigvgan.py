"""BigVGAN-v2 decode timing on the GPU box (CUDA events).  usage: python tools/bench_bigvgan.py [T=1250] [B=1] [iters=5]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "lemas-tts_b200")]
import torch

from lemas_tts import _native as nv
from lemas_tts import synthetic as syn
from lemas_tts.bigvgan import BigVGAN

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
arch = syn.FULL_BIGVGAN
voc = BigVGAN()
voc.load_state_dict(syn.make_bigvgan_state_dict(arch, seed=17), strict=True)
voc = voc.eval().to("cuda")
mel = syn.synthetic_ref_mel(B, T, 100, seed=5).permute(0, 2, 1).contiguous().cuda()
for _ in range(2):
    wav = voc(mel)
torch.cuda.synchronize()
l0 = nv.load().lemas_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    wav = voc(mel)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
launches = (nv.load().lemas_launch_count() - l0) // iters
audio_s = B * T * 256 / 24000
# algorithmic FLOPs of the generator (unpadded channels): convs of the AMP blocks + transposed convs + conv_pre
ch, L, flops = arch.upsample_initial_channel, T, 2.0 * T * 7 * 100 * arch.upsample_initial_channel
for r, k in zip(arch.upsample_rates, arch.upsample_kernel_sizes):
    flops += 2.0 * L * ch * (ch // 2) * k
    ch, L = ch // 2, L * r
    flops += 2.0 * L * ch * ch * sum(arch.resblock_kernel_sizes) * 6
flops *= B
print(f"BigVGAN-v2 decode B={B} T={T}: {ms:.2f} ms for {audio_s:.2f} s of audio ({audio_s / (ms * 1e-3):.0f}x real time), "
      f"{launches} launches, {flops / 1e12:.2f} TFLOP algorithmic -> {flops / (ms * 1e-3) / 1e12:.0f} TFLOP/s", flush=True)
