"""Full-size (BASELINE.json configs C2 / C5, 336 M-parameter model) checks on the GPU: one-step parity against the CPU
oracle at N = 2187, and size-independent properties at full step counts — graph replay == eager launches bit for
bit, identical utterances in a batch give identical rows, frames kept from the conditioning mel are bit-exact,
every value finite."""
import pytest
import torch

from lemas_tts import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full_model():
    from lemas_tts.model.backbones.dit import DiT
    from lemas_tts.model.cfm import CFM

    arch = syn.FULL_ARCH
    sd = syn.make_dit_state_dict(arch, seed=0)
    model = CFM(transformer=DiT(**arch.to_kwargs()), mel_spec_kwargs=dict(mel_spec_type="vocos"))
    model.load_state_dict(sd, strict=True)
    return model.cuda(), sd


def test_c2_one_step_matches_oracle(full_model):
    from oracle import lemas_oracle as orc

    model, sd = full_model
    cfg = syn.CONFIGS["C2"]
    arch = syn.FULL_ARCH
    cond = syn.synthetic_ref_mel(1, cfg.ref_frames, 100, seed=0)
    text = syn.synthetic_text_ids(1, cfg.n_text, 898, seed=0)
    noise = syn.synthetic_noise([cfg.total_frames], 100, seed=0)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    # one step over the whole interval: t = [0, 1] (any sway exponent leaves the end points in place), dt = 1
    ref_out, ref_traj = orc.cfm_sample(sd, arch, cond, text, cfg.total_frames, steps=1, cfg_strength=2.0,
                                       sway_sampling_coef=3.0, noise=noise, use_acc_grl=False)
    out, traj = model.sample(cond=cond.cuda(), text=text.cuda(), duration=cfg.total_frames, steps=1, cfg_strength=2.0,
                             sway_sampling_coef=3.0, noise=noise, use_acc_grl=False)
    got, ref = traj[-1].cpu(), ref_traj[-1]
    assert (ref - noise).abs().max() > 1.0, "the step must move the state for the comparison to mean anything"
    rel = ((got - ref).pow(2).mean() / ref.var()).item()
    print(f"C2 one full-size Euler step: mel-MSE/Var {rel:.3e}, max abs {(got - ref).abs().max():.3e}")
    assert rel <= 1e-5 and (got - ref).abs().max() <= 0.05


def test_c2_graph_equals_eager_and_keeps_reference(full_model):
    model, _ = full_model
    cfg = syn.CONFIGS["C2"]
    cond = syn.synthetic_ref_mel(1, cfg.ref_frames, 100, seed=0).cuda()
    text = syn.synthetic_text_ids(1, cfg.n_text, 898, seed=0).cuda()
    kw = dict(cond=cond, text=text, duration=cfg.total_frames, steps=cfg.steps, cfg_strength=cfg.cfg_strength,
              sway_sampling_coef=cfg.sway_coef, seed=7, use_acc_grl=False)
    out_g, _ = model.sample(**kw, return_trajectory=False)   # graph-replayed steps
    out_e, traj = model.sample(**kw, return_trajectory=True)  # eager launches (trajectory requested)
    assert torch.equal(out_g, out_e), "graph replay must not change a single bit"
    assert torch.isfinite(out_g).all() and traj.shape[0] == cfg.steps + 1
    assert torch.equal(out_g[:, :cfg.ref_frames], cond), "reference frames are copied, not computed"
    assert out_g[:, cfg.ref_frames:].abs().max() < 1e3


def test_identical_utterances_in_a_batch_give_identical_rows(full_model):
    model, _ = full_model
    N, Tc = 768, 256
    cond = syn.synthetic_ref_mel(1, Tc, 100, seed=2).repeat(3, 1, 1).cuda()
    text = syn.synthetic_text_ids(1, 120, 898, seed=2).repeat(3, 1).cuda()
    noise = syn.synthetic_noise([N], 100, seed=2).repeat(3, 1, 1)
    out, _ = model.sample(cond=cond, text=text, duration=N, steps=4, cfg_strength=2.0, sway_sampling_coef=3.0,
                          noise=noise, use_acc_grl=False, return_trajectory=False)
    assert torch.equal(out[0], out[1]) and torch.equal(out[0], out[2])


def test_c5_edit_keeps_unmasked_frames(full_model):
    model, _ = full_model
    cfg = syn.CONFIGS["C5"]
    cond = syn.synthetic_ref_mel(1, cfg.ref_frames, 100, seed=3).cuda()
    text = syn.synthetic_text_ids(1, cfg.n_text, 898, seed=3).cuda()
    edit = torch.ones(1, cfg.ref_frames, dtype=torch.bool, device="cuda")
    edit[:, 1125:1406] = False
    out, _ = model.sample(cond=cond, text=text, duration=cfg.ref_frames - 1, steps=6, cfg_strength=cfg.cfg_strength,
                          sway_sampling_coef=cfg.sway_coef, seed=3, edit_mask=edit, use_acc_grl=False,
                          return_trajectory=False)
    assert out.shape == (1, cfg.total_frames, 100)  # duration silently raised to lens + 1 (cfm.py:300)
    keep = edit[0]
    assert torch.equal(out[0, :cfg.ref_frames][keep], cond[0][keep])
    assert not torch.allclose(out[0, 1125:1406], cond[0, 1125:1406]), "masked span must be regenerated"
    assert torch.isfinite(out).all()
