"""training with the R diffusion copies of a sample packed behind ONE shared decoder prefix
(MLA.share_diffusion_prefix) gives the loss, the noise prediction and the gradients of the reference's repeated batch
(models/mla/model_mla.py:147-176) — image-only configuration, head_dim 128 (tcgen05 attention), one right-padded prompt."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _build(seed=0):
    from mla_b200.backbone import LLMBackbone, LlamaConfig
    from mla_b200.mla import MLA
    from mla_b200.vlm import PrismaticVLM
    from oracle import fixtures
    h, f, L, heads, T = 256, 512, 2, 2, 1
    flags = dict(use_diff=True, use_pointcloud=False, use_tactile=False, use_contrastive=False, use_generation=False)
    cfg = LlamaConfig(vocab_size=32064, hidden_size=h, intermediate_size=f, num_hidden_layers=L, num_attention_heads=heads)
    vlm = PrismaticVLM("tiny", LLMBackbone(config=cfg), token_size=h, action_dim=7, **flags)
    mla = MLA(vlm, None, token_size=h, action_dim=7, future_action_window_size=T, **flags)
    sd = fixtures.fill_state_dict(mla.state_dict(), seed=seed + 3)
    mla.load_state_dict({k: (v.to(torch.bfloat16).float() if torch.is_floating_point(v) else v) for k, v in sd.items()})
    mla = mla.cuda().train()
    mla.freeze_backbones("finetune")
    return mla, T


@pytest.mark.parametrize("pad_last,R", [(0, 4), (3, 3)])
def test_shared_prefix_matches_repeated_batch(cuda_lib, pad_last, R):
    from oracle import fixtures
    mla, T = _build()
    batch = fixtures.synthetic_batch(2, 8, T, 168, seed=77, pad_last=pad_last)
    kw = dict(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"], labels=batch["labels"],
              actions=batch["actions"], images=batch["images"], camera_name="rlbench_front", proprio=batch["proprio"],
              action_masks=batch["action_masks"], repeated_diffusion_steps=R, use_diff=True)
    names = ["vlm.llm_backbone.llm.model.layers.0.self_attn.q_proj.weight",
             "vlm.llm_backbone.llm.model.layers.0.self_attn.v_proj.weight",
             "vlm.llm_backbone.llm.model.layers.1.mlp.down_proj.weight",
             "vlm.llm_backbone.llm.model.layers.0.input_layernorm.weight", "vlm.llm_backbone.llm.model.norm.weight",
             "vlm.projector_2d.mlp.2.weight", "vlm.x_embedder.mlp.fc1.weight", "vlm.t_embedder.mlp.0.bias",
             "vlm.proprio_embedder.mlp.fc2.weight", "vlm.final_layer.mlp.fc2.weight",
             "vlm.llm_backbone.llm.model.embed_tokens.weight"]
    res = {}
    for share in (False, True):
        mla.share_diffusion_prefix = share
        for p in mla.parameters():
            p.grad = None
        mla.vlm.llm_backbone.llm.model.mark_grads_fresh()
        torch.manual_seed(5)
        loss_dict, out = mla(**kw)
        loss_dict["total_loss"].backward()
        mla.vlm.check_errors()
        named = dict(mla.named_parameters())
        res[share] = dict(loss=float(loss_dict["total_loss"]), noise_pred=out.noise_pred.detach().float().clone(),
                          noise=out.noise.clone(), t=out.timestep.clone(), lti=out.last_true_indices.clone(),
                          rows=out.hidden_states[0].shape, grads={k: named[k].grad.detach().float().clone() for k in names})
    a, b = res[False], res[True]
    assert torch.equal(a["noise"], b["noise"]) and torch.equal(a["t"], b["t"])          # same draws
    assert torch.equal(a["lti"].cpu(), b["lti"].cpu())
    # the decoder ran B * (F + Lt + R*(T+3)) rows instead of B * R * (F + Lt + T + 2)
    assert b["rows"][0] * b["rows"][1] < a["rows"][0] * a["rows"][1] / (R * 0.75)
    assert abs(a["loss"] - b["loss"]) <= 2e-3 * abs(a["loss"]), (a["loss"], b["loss"])
    assert rel_err(b["noise_pred"], a["noise_pred"]) < 1e-2
    for k in names:
        e = rel_err(b["grads"][k], a["grads"][k])
        assert e < 3e-2, (k, e)
        assert abs(float(b["grads"][k].norm()) - float(a["grads"][k].norm())) <= 1e-2 * float(a["grads"][k].norm()), k


def test_grouped_attention_kernels_match_dense_reference(cuda_lib):
    """The shared-prefix mask of the tcgen05 attention kernels against fp32 attention with an explicit mask: a row of a
    suffix group sees the prefix and, causally, its own group; forward and backward, with padding rows."""
    from mla_b200 import ops
    torch.manual_seed(3)
    B, S, H, D, n = 2, 300, 2, 128, 5
    P = torch.tensor([270, 151], dtype=torch.int32, device="cuda")       # second sequence: shorter prefix + filler rows
    groups = 6
    mask = torch.ones(B, S, dtype=torch.bool, device="cuda")
    mask[1, 151 + groups * n:] = False
    h = H * D
    qkv = (torch.randn(B * S, 3 * h, device="cuda") * 0.5).to(torch.bfloat16)
    dctx = (torch.randn(B * S, h, device="cuda") * 0.2).to(torch.bfloat16) * mask.view(-1, 1)
    ctx, lse = ops.attn_fwd(qkv, B, S, H, D, mask, grouped=(P, n))
    dqkv = ops.attn_bwd(dctx, qkv, ctx, lse, B, S, H, D, mask, grouped=(P, n))
    torch.cuda.synchronize()
    i = torch.arange(S, device="cuda")
    vis = torch.zeros(B, S, S, dtype=torch.bool, device="cuda")
    for b in range(B):
        p = int(P[b])
        causal = i[None, :] <= i[:, None]
        glo = torch.where(i >= p, p + ((i - p) // n) * n, torch.zeros_like(i))
        vis[b] = causal & ((i[None, :] < p) | (i[None, :] >= glo[:, None])) & mask[b][None, :] & mask[b][:, None]
    qf = qkv.float().requires_grad_(True)
    q, k, v = [qf[:, j * h:(j + 1) * h].reshape(B, S, H, D).transpose(1, 2) for j in range(3)]
    sc = (q @ k.transpose(-1, -2)) * D ** -0.5
    sc = sc.masked_fill(~vis[:, None], float("-inf"))
    pr = torch.softmax(sc, -1).nan_to_num(0.0)
    ref = (pr @ v).transpose(1, 2).reshape(B * S, h)
    ref.backward(dctx.float())
    valid = mask.view(-1)
    assert rel_err(ctx[valid], ref[valid]) < 6e-3
    assert ctx[~valid].abs().max() == 0
    for j in range(3):
        sl = slice(j * h, (j + 1) * h)
        assert rel_err(dqkv[:, sl][valid], qf.grad[:, sl][valid]) < 2e-2, j
