"""Inference denoise loop (MLA.predict_action_diff's DDIM loop, models/mla/model_mla.py:709-772) on the CUDA path:
decode kernels against fp32 references, the DDIM update bit-exact against the oracle (itself pinned bit-exact to the
reference's recorded steps), and the whole loop — reference schedule (full forward per step) and K/V-cached — against
the goldens recorded from the unmodified reference (tests/golden/ddim_*.npz) and the oracle's fp32 truth."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import rel_err
from test_oracle_vs_golden import oracle_cfg
from test_sampler_cpu import DDIM_CASES, load_ddim, state_dict_for, step_draws

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16)


@pytest.mark.parametrize("m", [1, 2, 3, 4, 8, 16, 17, 34])
@pytest.mark.parametrize("residual", [False, True])
def test_gemv_kernel(cuda_lib, m, residual):
    from mla_b200 import ops
    torch.manual_seed(m)
    for n, k in ((1536, 512), (384, 1376), (1027, 4096), (130, 11008)):
        x = _bf(torch.randn(m, k, device="cuda"))
        w = _bf(torch.randn(n, k, device="cuda") * k ** -0.5)
        r = _bf(torch.randn(m, n, device="cuda")) if residual else None
        want = x.float() @ w.float().t()
        if residual:
            want = _bf(want).float() + r.float()
        got = ops.gemv(x, w, residual=r)
        assert got.shape == (m, n) and got.dtype == torch.bfloat16
        assert rel_err(got, want) < 4e-3, rel_err(got, want)
    # a strided view as input (rows of a wider buffer), as the decode path passes it
    wide = _bf(torch.randn(m, 3 * 512, device="cuda"))
    w = _bf(torch.randn(256, 512, device="cuda") * 512 ** -0.5)
    assert rel_err(ops.gemv(wide[:, :512], w), wide[:, :512].float() @ w.float().t()) < 4e-3


@pytest.mark.parametrize("m", [1, 2, 3, 4, 6])
def test_gemv_fused_prologues(cuda_lib, m):
    """RMSNorm / SwiGLU applied while the activations are loaded == the separate kernels followed by the plain gemv
    (m <= 4 runs fused, m = 6 exercises the unfused fallback of the same call)."""
    from mla_b200 import ops
    torch.manual_seed(10 + m)
    for n, k in ((512, 128), (1024, 4096), (256, 11008)):
        x = _bf(torch.randn(m, k, device="cuda") * 1.7)
        lnw = _bf(1 + 0.1 * torch.randn(k, device="cuda"))
        w = _bf(torch.randn(n, k, device="cuda") * k ** -0.5)
        r = _bf(torch.randn(m, n, device="cuda"))
        want = ops.gemv(ops.rmsnorm_fwd(x, lnw, 1e-5), w, residual=r)
        got = ops.gemv(x, w, residual=r, norm=(lnw, 1e-5))
        assert rel_err(got, want) < 3e-3, ("norm", n, k, rel_err(got, want))
        gu = _bf(torch.randn(m, 2 * k, device="cuda"))
        want = ops.gemv(ops.swiglu_fwd(gu), w, residual=r)
        got = ops.gemv(gu, w, residual=r, swiglu=True)
        assert rel_err(got, want) < 1e-6, ("swiglu", n, k, rel_err(got, want))
        # and against fp32 math
        ref = torch.nn.functional.silu(gu[:, :k].float()) * gu[:, k:].float()
        assert rel_err(got, _bf(_bf(ref).float() @ w.float().t()).float() + r.float()) < 1e-2


def test_rope_cache_kernel(cuda_lib):
    """RoPE of the new rows + append to the K/V cache in one kernel == rope_ in place followed by the copy."""
    from mla_b200 import ops
    torch.manual_seed(4)
    for D, H in ((32, 4), (128, 8)):
        B, n, P = 2, 3, 11
        h = H * D
        qkv = _bf(torch.randn(B * n, 3 * h, device="cuda"))
        cache = _bf(torch.randn(B * (P + n), 2 * h, device="cuda"))
        pos = torch.arange(P + n, device="cuda").float()
        inv = 1.0 / (10000 ** (torch.arange(0, D, 2, device="cuda").float() / D))
        fr = pos[:, None] * inv[None]
        cos, sin = _bf(fr.cos())[P:].contiguous(), _bf(fr.sin())[P:].contiguous()
        q2, c2 = qkv.clone(), cache.clone()
        ops.rope_(q2, 0, 2 * H, D, n, cos, sin)
        c2.view(B, P + n, 2 * h)[:, P:] = q2.view(B, n, 3 * h)[:, :, h:]
        ops.rope_cache(qkv, cache, cos, sin, B, n, P, H, D)
        assert torch.equal(qkv[:, :h], q2[:, :h])            # q rotated in place
        assert torch.equal(cache, c2)                        # k rotated into the cache, v copied, prefix untouched


@pytest.mark.parametrize("P", [77, 300, 546])
@pytest.mark.parametrize("D,H", [(32, 4), (128, 8)])
def test_decode_attn_with_fused_rope(cuda_lib, D, H, P):
    """RoPE of q and of the new keys inside the attention kernel (they are read un-rotated from the projection and
    never appended to the cache) == rope_cache followed by decode_attn on the appended cache."""
    from mla_b200 import ops
    torch.manual_seed(D)
    B, n = 2, 3
    h = H * D
    qkv = _bf(torch.randn(B * n, 3 * h, device="cuda"))
    cache = _bf(torch.randn(B * (P + n), 2 * h, device="cuda"))
    pos = torch.arange(P + n, device="cuda").float()
    inv = 1.0 / (10000 ** (torch.arange(0, D, 2, device="cuda").float() / D))
    fr = pos[:, None] * inv[None]
    cos, sin = _bf(fr.cos())[P:].contiguous(), _bf(fr.sin())[P:].contiguous()
    # the denoise loop keeps the prefix head-major: [B, 2, H, P, D]
    hm = cache.view(B, P + n, 2, H, D)[:, :P].permute(0, 2, 3, 1, 4).contiguous()
    hm0 = hm.clone()
    got = ops.decode_attn_rope(qkv, hm, cos, sin, B, H, n, P, D, split_k=True)
    q2, c2 = qkv.clone(), cache.clone()
    ops.rope_cache(q2, c2, cos, sin, B, n, P, H, D)
    want = ops.decode_attn(q2, c2, B, H, n, P + n, D)
    # above 128 keys the kernel splits them over several CTAs and the last one merges: same result up to fp32 order
    one = ops.decode_attn_rope(qkv, hm, cos, sin, B, H, n, P, D, split_k=False)
    assert rel_err(one, want) < 1e-6, rel_err(one, want)
    assert rel_err(got, want) < (1e-6 if P + n <= 128 else 4e-3), rel_err(got, want)
    again = ops.decode_attn_rope(qkv, hm, cos, sin, B, H, n, P, D, split_k=True)           # counters re-armed
    assert torch.equal(again, got)
    assert torch.equal(hm, hm0)                                                             # cache untouched


@pytest.mark.parametrize("D", [32, 128])
def test_decode_attn_kernel(cuda_lib, D):
    """Suffix queries against the cache with flash-attn's bottom-right aligned causal mask."""
    from mla_b200 import ops
    torch.manual_seed(D)
    B, H, Lq, Lk = 2, 4, 3, 77
    q = _bf(torch.randn(B * Lq, 3 * H * D, device="cuda"))          # packed q|k|v rows: queries are the leading H*D
    kv = _bf(torch.randn(B * Lk, 2 * H * D, device="cuda"))
    got = ops.decode_attn(q, kv, B, H, Lq, Lk, D)
    qf = q[:, :H * D].float().view(B, Lq, H, D).transpose(1, 2)
    kf = kv[:, :H * D].float().view(B, Lk, H, D).transpose(1, 2)
    vf = kv[:, H * D:].float().view(B, Lk, H, D).transpose(1, 2)
    s = qf @ kf.transpose(-1, -2) * D ** -0.5
    i = torch.arange(Lq, device="cuda").view(Lq, 1)
    j = torch.arange(Lk, device="cuda").view(1, Lk)
    s = s.masked_fill(j > Lk - Lq + i, float("-inf"))
    want = (s.softmax(-1) @ vf).transpose(1, 2).reshape(B * Lq, H * D)
    assert rel_err(got, want) < 6e-3, rel_err(got, want)


@pytest.mark.parametrize("eps_dtype", [torch.bfloat16, torch.float32])
def test_ddim_step_bit_exact(cuda_lib, eps_dtype):
    from mla_b200 import ops
    from mla_b200.modules import create_diffusion
    from oracle import sampler as S
    torch.manual_seed(3)
    for n in (8, 4, 10):
        dd = create_diffusion("ddim%d" % n, "squaredcos_cap_v2", 100)
        keep, tab = S.ddim_tables(n)
        assert dd.timestep_map == keep and dd.num_timesteps == n
        coef = dd.coef("cuda")
        for i in range(n):
            x = torch.randn(5, 16, 7) * 3
            eps = torch.randn(5, 16, 7).to(eps_dtype)
            want = S.ddim_step(x, eps, i, tab)
            got = ops.ddim_step(x.cuda(), eps.cuda(), coef[i])
            assert torch.equal(got.cpu(), want), (n, i, (got.cpu() - want).abs().max())


def build_cuda_model(z, base):
    c, mla, sd = state_dict_for(z, base)
    mla.load_state_dict({k: (v.float() if torch.is_floating_point(v) else v) for k, v in sd.items()})
    return c, mla.cuda().eval(), sd


def to_cuda_batch(batch):
    b = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items() if k != "images"}
    b["images"] = {k: v.cuda() for k, v in batch["images"].items()}
    return b


@pytest.mark.parametrize("name", sorted(DDIM_CASES))
def test_eval_forward_matches_reference_per_step(cuda_lib, name):
    """PrismaticVLM.forward in eval mode at recorded DDIM steps (the reference's own x_t, FPS starts and neighbour
    sets): noise prediction vs the golden and vs the oracle's fp32 truth."""
    from mla_b200 import pointcloud_impl
    from oracle import mla as O
    z, batch = load_ddim(name)
    c, mla, sd = build_cuda_model(z, DDIM_CASES[name])
    sd32 = {k: (v.float() if torch.is_floating_point(v) else v) for k, v in sd.items()}
    cfg = dict(oracle_cfg(c), eval=True, repeated_diffusion_steps=1)
    cb = to_cuda_batch(batch)
    for s in (0, int(z["ddim_steps"]) - 1):
        x = torch.from_numpy(z[f"step{s}.x"])
        d = step_draws(z, s, x)
        pointcloud_impl.set_test_overrides(d.get("fps_starts"), d.get("knn_idx"))
        try:
            with torch.no_grad():
                _, eps = mla.vlm(x.cuda(), d["timestep"].cuda(), input_ids=cb["input_ids"], images=cb["images"],
                                 point_cloud=cb.get("point_cloud"), proprio=cb["proprio"], camera_name="rlbench_front")
                tru = O.forward(sd32, batch, cfg, d, compute_dtype=torch.float32)["noise_pred"]
                ref = O.forward(sd, batch, cfg, d, compute_dtype=torch.bfloat16, flavor="cuda")["noise_pred"]
        finally:
            pointcloud_impl.set_test_overrides(None, None)
        mla.vlm.check_errors()
        e_got, e_ref = rel_err(eps.float().cpu(), tru), rel_err(ref, tru)
        assert e_got < 1.5 * e_ref + 2e-3, (name, s, e_got, e_ref)
        assert rel_err(eps.float().cpu(), torch.from_numpy(z[f"step{s}.eps"])) < 3e-2


def test_denoise_loop_matches_reference_sample(cuda_lib):
    """The whole 8-step loop on the image-only Tiny-MLA: the reference's schedule (full forward per step) and the
    K/V-cached schedule both land on the reference's recorded sample, as close to the fp32 truth as its bf16 run."""
    from oracle import mla as O, sampler as S
    name = "ddim_tiny_img"
    z, batch = load_ddim(name)
    c, mla, sd = build_cuda_model(z, DDIM_CASES[name])
    cb = to_cuda_batch(batch)
    noise = torch.from_numpy(z["noise"]).cuda()
    n = int(z["ddim_steps"])
    full = mla.denoise_actions(cb["input_ids"], cb["images"], proprio=cb["proprio"], noise=noise, num_ddim_steps=n,
                               use_kv_cache=False)
    cached = mla.denoise_actions(cb["input_ids"], cb["images"], proprio=cb["proprio"], noise=noise, num_ddim_steps=n,
                                 use_kv_cache=True)
    # default = prefill + whole DDIM loop replayed as CUDA graphs; eager launches of the same kernels agree, and a
    # second replay (new noise, same session) is consistent with a fresh eager run
    eager = mla.denoise_actions(cb["input_ids"], cb["images"], proprio=cb["proprio"], noise=noise, num_ddim_steps=n,
                                use_kv_cache=True, use_cuda_graph=False)
    assert rel_err(cached, eager) < 1e-3, rel_err(cached, eager)
    noise2 = torch.randn_like(noise)
    g2 = mla.denoise_actions(cb["input_ids"], cb["images"], proprio=cb["proprio"], noise=noise2, num_ddim_steps=n)
    e2 = mla.denoise_actions(cb["input_ids"], cb["images"], proprio=cb["proprio"], noise=noise2, num_ddim_steps=n,
                             use_cuda_graph=False)
    assert rel_err(g2, e2) < 1e-3, rel_err(g2, e2)
    assert len(mla.vlm._denoise_sessions) == 1
    gold = torch.from_numpy(z["sample"])
    sd32 = {k: (v.float() if torch.is_floating_point(v) else v) for k, v in sd.items()}
    cfg = dict(oracle_cfg(c), eval=True, repeated_diffusion_steps=1)

    def model(x, t):
        with torch.no_grad():
            return O.forward(sd32, batch, cfg, dict(timestep=t, x=x), compute_dtype=torch.float32)["noise_pred"]
    tru = S.ddim_sample_loop(model, torch.from_numpy(z["noise"]), n)
    e_ref = rel_err(gold, tru)
    for tag, got in (("full", full), ("cached", cached)):
        assert got.shape == gold.shape and got.dtype == torch.float32
        e = rel_err(got.cpu(), tru)
        assert e < 1.5 * e_ref + 3e-3, (tag, "vs fp32 truth", e, "reference bf16 vs truth", e_ref)
        assert rel_err(got.cpu(), gold) < 3e-2, (tag, "vs golden", rel_err(got.cpu(), gold))
    assert rel_err(cached, full) < 2e-2, rel_err(cached, full)


def test_cached_equals_full_with_point_cloud(cuda_lib):
    """Point-cloud model (eval-mode BatchNorm on running statistics, T = 3 -> 4 action rows): with the FPS starts held
    fixed the cached schedule reproduces the full-forward schedule (the reference redraws the starts every step, which
    only re-samples an input-independent random choice)."""
    from mla_b200 import pointcloud_impl
    name = "ddim_tiny_pc"
    z, batch = load_ddim(name)
    c, mla, sd = build_cuda_model(z, DDIM_CASES[name])
    cb = to_cuda_batch(batch)
    noise = torch.from_numpy(z["noise"]).cuda()
    d = step_draws(z, 0, None)
    pointcloud_impl.set_test_overrides(d["fps_starts"], d["knn_idx"])
    try:
        kw = dict(point_cloud=cb["point_cloud"], proprio=cb["proprio"], noise=noise, num_ddim_steps=int(z["ddim_steps"]))
        full = mla.denoise_actions(cb["input_ids"], cb["images"], use_kv_cache=False, **kw)
        cached = mla.denoise_actions(cb["input_ids"], cb["images"], use_kv_cache=True, **kw)
    finally:
        pointcloud_impl.set_test_overrides(None, None)
    assert torch.isfinite(full).all() and torch.isfinite(cached).all()
    assert rel_err(cached, full) < 2e-2, rel_err(cached, full)
    # and it stays in the neighbourhood of the reference's sample (whose FPS starts differ per step)
    assert rel_err(full.cpu(), torch.from_numpy(z["sample"])) < 0.25


def test_ragged_tag_positions_are_rejected(cuda_lib):
    name = "ddim_tiny_img"
    z, batch = load_ddim(name)
    c, mla, sd = build_cuda_model(z, DDIM_CASES[name])
    cb = to_cuda_batch(batch)
    ids = cb["input_ids"].clone()
    ids[0, -1], ids[0, -2] = 5, 29871          # sample 0: tag one position earlier
    with pytest.raises(NotImplementedError):
        mla.denoise_actions(ids, cb["images"], proprio=cb["proprio"], num_ddim_steps=8, use_kv_cache=True)
    out = mla.denoise_actions(ids, cb["images"], proprio=cb["proprio"], num_ddim_steps=8, use_kv_cache=False)
    assert torch.isfinite(out).all()


def test_graph_session_follows_weight_updates(cuda_lib):
    """The CUDA-graph session reads weights through persistent bf16 copies: after the masters change (a training
    step), the next call refreshes them and the replay matches a fresh eager run."""
    name = "ddim_tiny_img"
    z, batch = load_ddim(name)
    c, mla, sd = build_cuda_model(z, DDIM_CASES[name])
    cb = to_cuda_batch(batch)
    noise = torch.from_numpy(z["noise"]).cuda()
    kw = dict(proprio=cb["proprio"], noise=noise, num_ddim_steps=8)
    before = mla.denoise_actions(cb["input_ids"], cb["images"], **kw)
    with torch.no_grad():
        for n_, p_ in mla.named_parameters():
            if n_.endswith(("q_proj.weight", "final_layer.mlp.fc1.weight", "x_embedder.mlp.fc2.weight")):
                p_.mul_(1.05)
    after_graph = mla.denoise_actions(cb["input_ids"], cb["images"], **kw)
    after_eager = mla.denoise_actions(cb["input_ids"], cb["images"], use_cuda_graph=False, **kw)
    assert rel_err(after_graph, after_eager) < 1e-3, rel_err(after_graph, after_eager)
    assert rel_err(after_graph, before) > 1e-3
