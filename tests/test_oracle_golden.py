"""Pins the CPU oracle (oracle/lemas_oracle.py) to golden vectors minted from the VERBATIM reference.

Generator: oracle/gen_golden.py (imports /root/reference/lemas_tts/model/{cfm,modules,backbones/dit}.py).
Bar: fp32 CPU vs fp32 CPU, same op order up to the restated split of third-party pieces -> 2e-5 abs.
"""
import pytest
import torch

import golden_cases as gc
from lemas_tts import synthetic as syn
from oracle import lemas_oracle as orc

TOL = 2e-5


@pytest.fixture(scope="module", autouse=True)
def _threads():
    torch.set_num_threads(8)


@pytest.mark.parametrize("name", list(gc.CASES))
def test_sample_matches_reference(name):
    case = gc.CASES[name]
    gold = gc.load(name)
    inp = gc.inputs(case)
    sums = torch.tensor([inp["cond"].double().abs().sum(), inp["noise"].double().abs().sum(),
                         float(inp["text"].sum())], dtype=torch.float64)
    assert torch.allclose(sums, gold["in_sums"], rtol=0, atol=1e-6), "synthetic inputs drifted from the fixture"

    sd = syn.make_dit_state_dict(inp["arch"], seed=case["wseed"])
    out, traj = orc.cfm_sample(sd, inp["arch"], inp["cond"], inp["text"], inp["duration"], lens=inp["lens"],
                               steps=case["steps"], cfg_strength=case["cfg"], sway_sampling_coef=case["sway"],
                               noise=inp["noise"], edit_mask=inp["edit_mask"], use_acc_grl=case["use_acc_grl"])
    assert out.shape == gold["out"].shape
    assert (traj[1] - gold["first_step"]).abs().max() < TOL
    assert (traj[-1] - gold["last"]).abs().max() < TOL * 5
    assert (out - gold["out"]).abs().max() < TOL * 5


@pytest.mark.parametrize("name", ["sample_tiny_b3_ragged", "sample_tiny_edit"])
def test_backbone_forward_matches_reference(name):
    case = gc.CASES[name]
    gold = gc.load(name)
    inp = gc.inputs(case)
    arch = inp["arch"]
    sd = syn.make_dit_state_dict(arch, seed=case["wseed"])
    N = gold["last"].shape[1]
    mask = None
    if case["batch"] > 1:
        mask = torch.arange(N)[None] < torch.tensor(inp["durations"])[:, None]
    cond = torch.nn.functional.pad(inp["cond"], (0, 0, 0, N - inp["cond"].shape[1]))
    t = torch.tensor(0.37)
    for drop, key in ((False, "fwd_cond"), (True, "fwd_uncond")):
        te = orc.text_embedding(sd, arch, inp["text"], N, drop_text=drop)
        got = orc.dit_forward(sd, arch, gold["last"], cond, te, t, mask, drop_audio_cond=drop)
        assert (got - gold[key]).abs().max() < TOL


def test_time_grids_match_reference():
    grids = torch.load(gc.GOLDEN / "time_grids.pt", weights_only=True)
    for key, want in grids.items():
        steps, coef = key.split("_")
        coef = None if coef == "None" else float(coef)
        got = orc.time_grid(int(steps), coef)
        assert torch.equal(got, want), key


def test_melspec_matches_reference():
    want = torch.load(gc.GOLDEN / "melspec.pt", weights_only=True)["mel"]
    got = orc.mel_spectrogram(syn.synthetic_ref_audio(2, 24000, seed=9))
    assert (got - want).abs().max() < 1e-5


def test_edit_mask_keeps_reference_region_bit_exact():
    case = gc.CASES["sample_tiny_edit"]
    inp = gc.inputs(case)
    gold = gc.load("sample_tiny_edit")
    keep = torch.ones(150, dtype=torch.bool)
    keep[case["edit"][0]: case["edit"][1]] = False
    assert torch.equal(gold["out"][0, :150][keep], inp["cond"][0][keep])


def test_prosody_conditioning_matches_reference():
    """Config C3's path (raw-audio cond, prosody encoder on): the oracle's conditioning algebra (prosody_to_mel added
    after zero padding, use_acc_grl taking the pre-prosody mel, prosody_text_proj on both CFG variants) against the
    verbatim reference; the embeddings themselves come from the fixture (the ECAPA-TDNN is pinned separately in
    tests/test_host_cpu.py)."""
    case = gc.PROSODY_CASE
    gold = gc.load(case["name"])
    arch, audio, text, noise, sd = gc.prosody_inputs(case)
    mel = orc.mel_spectrogram(audio).permute(0, 2, 1)
    for grl in (False, True):
        out, traj = orc.cfm_sample(sd, arch, mel, text, torch.tensor(case["durations"]), lens=torch.tensor(case["lens"]),
                                   steps=case["steps"], cfg_strength=case["cfg"], sway_sampling_coef=case["sway"],
                                   noise=noise, use_acc_grl=grl, prosody_embeds=gold["embeds"])
        assert (traj[-1] - gold[f"last_grl{int(grl)}"]).abs().max() < TOL * 5
        assert (out - gold[f"out_grl{int(grl)}"]).abs().max() < TOL * 5
    assert (gold["out_grl0"] - gold["out_grl1"]).abs().max() > 1e-2, "the two settings must differ for the test to bite"


def test_rotary_pair_convention_matches_an_independent_implementation():
    """x-transformers (pinned >= 1.31.14, absent offline) rotates ADJACENT pairs (x0, x1) -> (x0 cos - x1 sin, x1 cos +
    x0 sin) with angle n * 10000^(-2j/64) shared by columns 2j, 2j+1 (call sites: /root/reference/lemas_tts/model/
    backbones/dit.py:236, modules.py:476-480).  GPT-J in `transformers` (installed) is an independent implementation of
    exactly that convention (`rotate_every_two`, interleaved sin / cos): the oracle's table + rotation and the shim the
    verbatim reference runs on (oracle/verbatim.py) must agree with it bit for bit in fp32."""
    gptj = pytest.importorskip("transformers.models.gptj.modeling_gptj")
    from oracle import lemas_oracle as orc
    from oracle import verbatim

    N, H, dh = 257, 4, 64
    g = torch.Generator().manual_seed(5)
    t = torch.randn(2, H, N, dh, generator=g)
    inv_freq = 1.0 / (10000 ** (torch.arange(0, dh, 2).float() / dh))
    ang = orc.rotary_table(inv_freq, N)
    got = orc.apply_rotary(t, ang)
    # GPT-J layout: [batch, seq, heads, dim]; sin / cos [1, seq, dim/2]
    half = torch.outer(torch.arange(N).float(), inv_freq)
    want = gptj.apply_rotary_pos_emb(t.transpose(1, 2), half.sin()[None], half.cos()[None]).transpose(1, 2)
    assert torch.equal(got, want)
    rot = verbatim._Rotary(dh)
    freqs, scale = rot.forward_from_seq_len(N)
    assert torch.equal(verbatim._apply_rotary(t, freqs, scale), want)
