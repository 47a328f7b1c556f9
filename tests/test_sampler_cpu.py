"""Pins the oracle's DDIM sampler (oracle/sampler.py) and eval-mode forward (oracle/mla.py, cfg eval=True) to the
reference's inference denoise loop: tests/golden/ddim_*.npz were recorded by tests/golden/make_golden_ddim.py from the
UNMODIFIED reference (MLA.create_ddim + ddim_diffusion.ddim_sample_loop over PrismaticVLM.forward in eval mode)."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from test_oracle_vs_golden import build_state_dict, case_cfg, load_case, oracle_cfg

DDIM_CASES = {"ddim_tiny_img": "tiny_img", "ddim_tiny_pc": "tiny_pc"}


def load_ddim(name):
    z, _ = load_case(name)
    batch = {"input_ids": torch.from_numpy(z["input_ids"]), "proprio": torch.from_numpy(z["proprio"]),
             "images": {"front_image": torch.from_numpy(z["front_image"])}, "attention_mask": None}
    if "point_cloud" in z.files:
        batch["point_cloud"] = torch.from_numpy(z["point_cloud"])
    return z, batch


def state_dict_for(z, base):
    """Same deterministic weights as the golden run, plus the BatchNorm running statistics it recorded."""
    c = case_cfg(base)
    mla, sd = build_state_dict(c)
    for k in z.files:
        if k.startswith("buf."):
            assert k[4:] in sd, k
            sd[k[4:]] = torch.from_numpy(z[k])
    return c, mla, sd


def step_draws(z, s, x):
    d = dict(timestep=torch.from_numpy(z[f"step{s}.t"]), x=x)
    starts = [torch.from_numpy(z[k]) for k in sorted(f for f in z.files if f.startswith(f"step{s}.fps_start_"))]
    if starts:
        d["fps_starts"] = starts
        d["knn_idx"] = [torch.from_numpy(z[k].astype(np.int64))
                        for k in sorted(f for f in z.files if f.startswith(f"step{s}.knn_idx_"))]
    return d


@pytest.mark.parametrize("name", sorted(DDIM_CASES))
def test_schedule_and_update_match_reference(name):
    from oracle import sampler as S
    z, _ = load_ddim(name)
    n = int(z["ddim_steps"])
    keep, ac = S.ddim_schedule(n)
    assert keep == z["timestep_map"].tolist()
    assert np.array_equal(ac, z["alphas_cumprod"])                       # float64, bit-exact
    _, tab = S.ddim_tables(n)
    for s in range(n):                                                  # recorded order: i = n-1 .. 0
        i = n - 1 - s
        assert int(z[f"step{s}.t"][0]) == keep[i]
        nxt = torch.from_numpy(z[f"step{s + 1}.x"]) if s + 1 < n else torch.from_numpy(z["sample"])
        got = S.ddim_step(torch.from_numpy(z[f"step{s}.x"]), torch.from_numpy(z[f"step{s}.eps"]), i, tab)
        assert torch.equal(got, nxt), (s, (got - nxt).abs().max())
    assert torch.equal(torch.from_numpy(z["step0.x"]), torch.from_numpy(z["noise"]))


@pytest.mark.parametrize("name", sorted(DDIM_CASES))
def test_oracle_eval_forward_reproduces_reference_noise_prediction(name):
    """PrismaticVLM.forward in eval mode (tag 29871, BatchNorm on running statistics, no attention mask) at every
    recorded DDIM step, fed the reference's own x_t: the oracle's bf16 replay matches the reference's prediction."""
    from oracle import mla as O
    z, batch = load_ddim(name)
    c, _, sd = state_dict_for(z, DDIM_CASES[name])
    cfg = dict(oracle_cfg(c), eval=True, repeated_diffusion_steps=1)
    for s in (0, int(z["ddim_steps"]) - 1):
        with torch.no_grad():
            out = O.forward(sd, batch, cfg, step_draws(z, s, torch.from_numpy(z[f"step{s}.x"])),
                            compute_dtype=torch.bfloat16, flavor="cpu")
        e = rel_err(out["noise_pred"], torch.from_numpy(z[f"step{s}.eps"]))
        assert e < 2e-2, (name, s, e)


def test_oracle_sampler_end_to_end_close_to_reference():
    """Whole loop through the oracle in fp32 (truth): lands within bf16 noise of the reference's bf16 sample."""
    from oracle import mla as O, sampler as S
    name = "ddim_tiny_img"
    z, batch = load_ddim(name)
    c, _, sd = state_dict_for(z, DDIM_CASES[name])
    sd32 = {k: (v.float() if torch.is_floating_point(v) else v) for k, v in sd.items()}
    cfg = dict(oracle_cfg(c), eval=True, repeated_diffusion_steps=1)

    def model(x, t):
        with torch.no_grad():
            return O.forward(sd32, batch, cfg, dict(timestep=t, x=x), compute_dtype=torch.float32)["noise_pred"]
    got = S.ddim_sample_loop(model, torch.from_numpy(z["noise"]), int(z["ddim_steps"]))
    assert rel_err(got, torch.from_numpy(z["sample"])) < 3e-2


@pytest.mark.parametrize("n", [1, 4, 8, 10, 20])
def test_product_schedule_equals_oracle(n):
    """Host logic of the CUDA sampler (mla_b200/sampler.py, no kernel runs): timestep map, float64 alphas and the
    four fp32 coefficients per step that mla_ddim_step consumes."""
    from mla_b200 import create_diffusion
    from oracle import sampler as S
    dd = create_diffusion("ddim%d" % n, "squaredcos_cap_v2", 100, sigma_small=True, learn_sigma=False)
    keep, tab = S.ddim_tables(n)
    assert dd.timestep_map == keep and dd.num_timesteps == n and dd.original_num_steps == 100
    assert np.array_equal(dd.alphas_cumprod, S.ddim_schedule(n)[1])
    want = np.stack([tab[:, 0].astype(np.float32), tab[:, 1].astype(np.float32),
                     np.sqrt(tab[:, 2].astype(np.float32)), np.sqrt(np.float32(1) - tab[:, 2].astype(np.float32))], 1)
    assert dd._coef.dtype == np.float32 and np.array_equal(dd._coef, want)
    with pytest.raises(NotImplementedError):
        dd.ddim_sample_loop(lambda x, t: x, (1, 1, 7), torch.zeros(1, 1, 7), eta=0.5)


def test_training_diffusion_unchanged_by_respacing_support():
    from mla_b200 import create_diffusion
    d = create_diffusion(timestep_respacing="", noise_schedule="squaredcos_cap_v2", diffusion_steps=100)
    assert d.num_timesteps == 100 and not hasattr(d, "timestep_map")
    with pytest.raises(NotImplementedError):
        create_diffusion(timestep_respacing="10,10", diffusion_steps=100)


def test_prompt_builder_matches_reference_when_available():
    """The one-turn prompt of predict_action_diff (model_mla.py:627-631) through our PurePromptBuilder vs the
    reference's (models/backbones/llm/prompting/base_prompter.py) — only where /root/reference exists."""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    ref_shim.load()
    from models.backbones.llm.prompting import PurePromptBuilder as Ref
    from mla_b200.backbone import PurePromptBuilder as Ours
    for text in ("close the jar", "  put <image> the rubbish in bin ", "stack 2 blocks"):
        msg = f"What action should the robot take to {text.lower()}?"
        a, b = Ref("prismatic"), Ours("prismatic")
        assert a.add_turn("human", msg) == b.add_turn("human", msg)
        assert a.get_prompt() == b.get_prompt()
        assert a.add_turn("gpt", "ok") == b.add_turn("gpt", "ok")
        assert a.get_prompt() == b.get_prompt()
        assert a.get_potential_prompt("next") == b.get_potential_prompt("next")
