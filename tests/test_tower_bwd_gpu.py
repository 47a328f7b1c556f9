"""Stage "pretrain" (models/vlm/prismatic.py:427-434 trains vision_tower_2d / vision_tower_3d): backward of the two
encoder-free tokenizers on the CUDA path.

(1) kernel-level: each kernel of csrc/tower_bwd.cu against fp32 autograd of the same formula (the reference ops are
    plain PyTorch: LocalAttention vision_tokenizer.py:27-47, nn.LayerNorm, train-mode nn.BatchNorm2d + ReLU/residual
    Point_PN.py:188-219, x.max(-1) :157, index_points :41-58);
(2) end to end: every tokenizer parameter's gradient of MLA.forward's total loss against the oracle's autograd, in
    fp32 (truth) and in the reference's bf16 arithmetic: ours must be as close to the truth as the reference
    arithmetic is (x1.5 + floor), like the forward tensors of test_mla_gpu.py.
"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err
from test_mla_gpu import _Draws
from test_oracle_vs_golden import build_state_dict, case_cfg, draws_of, load_case, oracle_cfg

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16)


# ------------------------------------------------------------------------------------------------ kernel level
def test_local_attn_bwd_kernel(cuda_lib):
    from mla_b200 import _lib, ops
    torch.manual_seed(0)
    G, Cc, heads, win = 37, 256, 8, 9
    scale = Cc ** -0.5
    q = _bf(torch.randn(G, Cc, device="cuda") * 2)
    kv = _bf(torch.randn(G * win, 2 * Cc, device="cuda"))
    do = _bf(torch.randn(G, Cc, device="cuda"))
    q32, kv32 = q.float().requires_grad_(True), kv.float().requires_grad_(True)
    dh = Cc // heads
    k_, v_ = kv32.view(G, win, 2, heads, dh).unbind(2)                    # [G, win, heads, dh]
    a = ((q32.view(G, 1, heads, dh) * scale) * k_).sum(-1)                # [G, win, heads]
    p = a.softmax(dim=1)
    out = (p.unsqueeze(-1) * v_).sum(1).reshape(G, Cc)
    out.backward(do.float())
    # forward parity first (same recompute is used by the backward)
    agg = torch.empty_like(q)
    _lib.check(cuda_lib.mla_local_attn(ops._p(q), ops._p(kv), ops._p(agg), C.c_int64(G), C.c_int32(Cc), C.c_int32(heads),
                                       C.c_int32(win), C.c_float(scale), ops._stream()))
    assert rel_err(agg, out.detach()) < 1e-2
    dq, dkv = torch.empty_like(q), torch.empty_like(kv)
    _lib.check(cuda_lib.mla_local_attn_bwd(ops._p(q), ops._p(kv), ops._p(do), ops._p(dq), ops._p(dkv), C.c_int64(G),
                                           C.c_int32(Cc), C.c_int32(heads), C.c_int32(win), C.c_float(scale),
                                           ops._stream()))
    assert rel_err(dq, q32.grad) < 3e-2, rel_err(dq, q32.grad)
    assert rel_err(dkv[:, :Cc], kv32.grad[:, :Cc]) < 3e-2, rel_err(dkv[:, :Cc], kv32.grad[:, :Cc])
    assert rel_err(dkv[:, Cc:], kv32.grad[:, Cc:]) < 1e-2, rel_err(dkv[:, Cc:], kv32.grad[:, Cc:])


@pytest.mark.parametrize("extras", [False, True])
def test_layernorm_bwd_kernel(cuda_lib, extras):
    from mla_b200 import ops
    from mla_b200.vision import layernorm_bwd
    torch.manual_seed(1)
    win, groups, h = 9, 23, 1024
    rows = groups * win
    x = _bf(torch.randn(rows, h, device="cuda") * 2 + 0.3)
    w = torch.randn(h, device="cuda")
    b = torch.randn(h, device="cuda")
    dy = _bf(torch.randn(rows, h, device="cuda"))
    dres = _bf(torch.randn(rows, h, device="cuda")) if extras else None
    dgrp = _bf(torch.randn(groups, h, device="cuda")) if extras else None
    x32 = x.float().requires_grad_(True)
    w32, b32 = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = F.layer_norm(x32, (h,), w32, b32, 1e-5)
    y.backward(dy.float())
    want = x32.grad
    if extras:
        want = want + dres.float() + dgrp.float().repeat_interleave(win, 0) / win
    # forward statistics must be the ones layernorm_fwd used
    assert rel_err(ops.layernorm(x, w, b, 1e-5), y.detach()) < 5e-3
    dx, dw, db = layernorm_bwd(dy, x, w, 1e-5, dres=dres, dgrp=dgrp, win=win)
    assert rel_err(dx, want) < 6e-3, rel_err(dx, want)
    assert rel_err(dw, w32.grad) < 1e-4, rel_err(dw, w32.grad)
    assert rel_err(db, b32.grad) < 1e-4, rel_err(db, b32.grad)


@pytest.mark.parametrize("mode,up_f32", [(0, False), (1, False), (1, True), (2, True)])
def test_bn_bwd_kernel(cuda_lib, mode, up_f32):
    """Train-mode BatchNorm over rows, masks as in Linear1Layer / Linear2Layer (Point_PN.py:173-219)."""
    from mla_b200 import _lib, ops
    from mla_b200.pointcloud_impl import _bn_bwd
    torch.manual_seed(2 + mode)
    groups, K, Cc = 64, 9, 96
    rows = groups * K
    y = _bf(torch.randn(rows, Cc, device="cuda") * 1.5 + 0.2)
    bn = torch.nn.BatchNorm1d(Cc).cuda().train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.5, 0.5)
    xres = torch.randn(rows, Cc, device="cuda")
    up = torch.randn(rows, Cc, device="cuda")
    if not up_f32:
        up = _bf(up)
    # statistics through our forward kernels
    sums = torch.zeros(2 * Cc, device="cuda")
    coef = torch.empty(2 * Cc, device="cuda")
    _lib.check(cuda_lib.mla_bn_stats(ops._p(y), ops._p(sums), C.c_int64(rows), C.c_int32(Cc), ops._stream()))
    _lib.check(cuda_lib.mla_bn_finalize(ops._p(sums), ops._p(coef), None, None, C.c_int64(rows), C.c_int32(Cc),
                                        C.c_float(bn.eps), C.c_float(0.1), ops._stream()))
    # fp32 autograd of the same op
    y32 = y.float().requires_grad_(True)
    w32, b32 = bn.weight.detach().clone().requires_grad_(True), bn.bias.detach().clone().requires_grad_(True)
    o = F.batch_norm(y32, None, None, w32, b32, True, 0.1, bn.eps)
    xnew = None
    if mode == 0:
        out = o
    elif mode == 1:
        out = F.relu(o)
    else:
        x32 = xres.clone().requires_grad_(True)
        out = F.relu(o + x32)
        xnew = torch.empty(rows, Cc, device="cuda")
        _lib.check(cuda_lib.mla_bn_res_relu(ops._p(y), ops._p(coef), ops._p(bn.weight), ops._p(bn.bias), ops._p(xres),
                                            ops._p(xnew), None, None, C.c_int64(groups), C.c_int32(K), C.c_int32(Cc),
                                            ops._stream()))
        assert rel_err(xnew, out.detach()) < 5e-3
    out.backward(up.float())
    dy, dw, db, dres = _bn_bwd(up, y, coef, bn, mode, xnew=xnew, want_res=mode == 2)
    tol = 2e-2 if mode == 2 else 8e-3        # mode 2 rounds the upstream to bf16 (the reference's bn output dtype)
    assert rel_err(dy, y32.grad) < tol, rel_err(dy, y32.grad)
    assert rel_err(dw, w32.grad) < tol, rel_err(dw, w32.grad)
    assert rel_err(db, b32.grad) < tol, rel_err(db, b32.grad)
    if mode == 2:
        # the forward rounds bn(y) to bf16 before the residual add (the reference's BatchNorm output dtype), so the
        # ReLU mask may differ from the fp32 graph's where the sum is within rounding of zero: compare away from it
        clear = out.detach().abs() > 2e-2
        assert clear.float().mean().item() > 0.45
        assert rel_err(dres[clear], x32.grad[clear]) < 1e-6, rel_err(dres[clear], x32.grad[clear])
        assert rel_err(dres, x32.grad) < 5e-2, rel_err(dres, x32.grad)


def test_maxpool_and_gather_bwd_kernels(cuda_lib):
    from mla_b200 import _lib, ops
    torch.manual_seed(5)
    B, N, G, K, Cc = 3, 40, 20, 7, 24
    xnew = torch.randn(B * G, K, 2 * Cc, device="cuda").relu()
    xr = xnew.clone().requires_grad_(True)
    dpool = torch.randn(B * G, 2 * Cc, device="cuda")
    xr.max(1)[0].backward(dpool)
    dx = torch.empty_like(xnew)
    _lib.check(cuda_lib.mla_maxpool_bwd(ops._p(xnew), ops._p(dpool), ops._p(dx), C.c_int64(B * G), C.c_int32(K),
                                        C.c_int32(2 * Cc), ops._stream()))
    live = (xnew.max(1, keepdim=True)[0] > 0).expand_as(xnew)      # all-zero columns tie; their gradient is masked later
    assert torch.equal(dx[live], xr.grad[live])
    # neighbour / centre gathers
    feat = torch.randn(B, N, Cc, device="cuda", requires_grad=True)
    fps_idx = torch.stack([torch.randperm(N, device="cuda")[:G] for _ in range(B)]).int()
    knn_idx = torch.randint(0, N, (B, G, K), device="cuda").int()
    bi = torch.arange(B, device="cuda").view(B, 1, 1)
    knn_x = feat[bi, knn_idx.long()]                                                         # [B,G,K,C]
    lc_x = feat[bi[:, :, 0], fps_idx.long()]                                                 # [B,G,C]
    X = torch.cat([knn_x, lc_x.unsqueeze(2).expand(-1, -1, K, -1)], -1)
    dX = torch.randn_like(X)
    X.backward(dX)
    dfeat = torch.zeros(B * N, Cc, device="cuda")
    _lib.check(cuda_lib.mla_group_pose_bwd(ops._p(dX.contiguous()), ops._p(fps_idx), ops._p(knn_idx), ops._p(dfeat),
                                           C.c_int32(B), C.c_int32(N), C.c_int32(G), C.c_int32(K), C.c_int32(Cc),
                                           ops._stream()))
    assert rel_err(dfeat.view(B, N, Cc), feat.grad) < 1e-5


def test_long_reduction_wgrad(cuda_lib):
    """dW of a 1x1 conv over ~1e5 neighbour rows: the block-expanded GEMM + diagonal fold equals the plain product."""
    from mla_b200.pointcloud_impl import _wgrad
    torch.manual_seed(6)
    rows, m, n = 73728, 96, 192
    dy = _bf(torch.randn(rows, m, device="cuda"))
    x = _bf(torch.randn(rows, n, device="cuda"))
    want = dy.float().t() @ x.float()
    got = _wgrad(dy, x)
    assert got.shape == (m, n)
    assert rel_err(got, want) < 1e-4, rel_err(got, want)
    small = _wgrad(dy[:1000], x[:1000])
    assert rel_err(small, dy[:1000].float().t() @ x[:1000].float()) < 1e-4


def test_nearest_center_kernel(cuda_lib):
    """Tactile positives (prismatic.py:742-749): torch.cdist + topk(1) + gather of the projected patch index."""
    from mla_b200 import _lib, ops
    torch.manual_seed(8)
    B, A, G, pw = 5, 2, 256, 16
    grip = torch.rand(B, A, 3, device="cuda")
    centers = torch.rand(B, G, 3, device="cuda")
    patch_idx = torch.randint(0, pw, (B, G, 2), device="cuda")
    pos = torch.empty((B, A, 1), dtype=torch.long, device="cuda")
    lin = torch.empty((B, A, 1), dtype=torch.long, device="cuda")
    _lib.check(cuda_lib.mla_nearest_center(ops._p(grip), ops._p(centers), ops._p(patch_idx), C.c_int32(B), C.c_int32(A),
                                           C.c_int32(G), C.c_int32(pw), ops._p(pos), ops._p(lin), ops._stream()))
    d = torch.cdist(grip, centers)
    ref = torch.topk(d, k=1, dim=2, largest=False)[1]
    picked = torch.gather(d, 2, pos)
    assert bool((picked <= d.min(2, keepdim=True)[0] * (1 + 1e-5)).all())
    assert (pos == ref).float().mean().item() >= 0.9
    sel = torch.gather(patch_idx.unsqueeze(1).expand(-1, A, -1, -1), 2, pos.unsqueeze(-1).expand(-1, -1, -1, 2))
    assert torch.equal(lin, sel[..., 0] * pw + sel[..., 1])


# ------------------------------------------------------------------------------------------------ end to end
@pytest.mark.parametrize("name", ["tiny_img", "tiny_pc", "align"])
def test_pretrain_stage_tokenizer_gradients(cuda_lib, name):
    from mla_b200 import pointcloud_impl
    from oracle import mla as O
    z, batch = load_case(name)
    c = case_cfg(name)
    mla, sd = build_state_dict(c, dtype=torch.bfloat16)
    mla.load_state_dict({k: (v.float() if torch.is_floating_point(v) else v) for k, v in sd.items()})
    mla = mla.cuda().train()
    mla.freeze_backbones("pretrain")
    assert mla.vlm.vision_backbone_requires_grad
    d = draws_of(z)
    pointcloud_impl.set_test_overrides(d.get("fps_starts"), d.get("knn_idx"))
    try:
        with _Draws(z):
            loss_dict, out = mla(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"],
                                 labels=batch["labels"], actions=batch["actions"], images=batch["images"],
                                 camera_name="rlbench_front", point_cloud=batch.get("point_cloud"),
                                 tactile=batch.get("tactile"), proprio=batch["proprio"],
                                 gripper_xyz=batch.get("gripper_xyz"), action_masks=batch["action_masks"],
                                 repeated_diffusion_steps=c["R"], use_diff=True)
        loss_dict["total_loss"].backward()
    finally:
        pointcloud_impl.set_test_overrides(None, None)
    mla.vlm.check_errors()
    # the forward is the frozen path's forward: same loss as the golden reference run
    assert abs(float(loss_dict["total_loss"]) - float(z["total_loss"])) <= 4e-3 * abs(float(z["total_loss"]))

    keys = [k for k in sd if k.startswith(("vlm.vision_tower_2d.", "vlm.vision_tower_3d."))
            and torch.is_floating_point(sd[k]) and "running_" not in k]
    sd32 = {k: (v.float() if torch.is_floating_point(v) else v) for k, v in sd.items()}
    sd16 = dict(sd)
    for k in keys:
        sd32[k] = sd32[k].clone().requires_grad_(True)
        sd16[k] = sd16[k].clone().requires_grad_(True)
    O.forward(sd32, batch, oracle_cfg(c), d, compute_dtype=torch.float32)["total_loss"].backward()
    O.forward(sd16, batch, oracle_cfg(c), d, compute_dtype=torch.bfloat16, flavor="cuda")["total_loss"].backward()

    named = dict(mla.named_parameters())
    gmax = max(float(sd32[k].grad.norm()) for k in keys if sd32[k].grad is not None)
    report, bad = [], []
    for k in keys:
        t, r, g = sd32[k].grad, sd16[k].grad, named[k].grad
        if t is None:                      # class/split embeddings, GlobalAttention, cls_token, pos_embed: unused
            assert g is None or float(g.abs().max()) == 0.0, k
            continue
        assert g is not None, k
        assert g.shape == t.shape and g.dtype == torch.float32, (k, g.shape, g.dtype)
        if float(t.norm()) < 1e-4 * gmax:  # conv bias in front of a train-mode BatchNorm: gradient is identically 0
            assert float(g.norm()) < 2e-3 * gmax, (k, float(g.norm()), gmax)
            continue
        e_got, e_ref = rel_err(g.cpu(), t), rel_err(r.float(), t)
        report.append((k, round(e_got, 4), round(e_ref, 4)))
        if not e_got < 1.5 * e_ref + 2e-2:
            bad.append((k, e_got, e_ref))
    assert not bad, (bad, report)


def test_pretrain_stage_trainer_step(cuda_lib):
    """One optimizer step with the tokenizers trainable: every tokenizer parameter that has a gradient moves, the
    unused ones (GlobalAttention, class/split embeddings, cls_token, pos_embed) stay put, BN running stats update."""
    from mla_b200 import pointcloud_impl
    from mla_b200.trainer import DataParallelTrainer
    name = "tiny_pc"
    z, batch = load_case(name)
    c = case_cfg(name)
    mla, sd = build_state_dict(c, dtype=torch.bfloat16)
    mla.load_state_dict({k: (v.float() if torch.is_floating_point(v) else v) for k, v in sd.items()})
    mla = mla.cuda().train()
    mla.freeze_backbones("pretrain")
    trainer = DataParallelTrainer(mla, lr=1e-3, weight_decay=0.0, max_grad_norm=1.0)
    before = {k: v.detach().clone() for k, v in mla.state_dict().items()}
    d = draws_of(z)
    pointcloud_impl.set_test_overrides(d.get("fps_starts"), d.get("knn_idx"))
    try:
        with _Draws(z):
            loss_dict, _ = mla(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"],
                               labels=batch["labels"], actions=batch["actions"], images=batch["images"],
                               camera_name="rlbench_front", point_cloud=batch.get("point_cloud"),
                               proprio=batch["proprio"], action_masks=batch["action_masks"],
                               repeated_diffusion_steps=c["R"], use_diff=True)
        loss_dict["total_loss"].backward()
    finally:
        pointcloud_impl.set_test_overrides(None, None)
    trainer.step()
    after = mla.state_dict()
    moved = lambda k: not torch.equal(before[k], after[k])
    assert moved("vlm.vision_tower_2d.patch_embedding.weight")
    assert moved("vlm.vision_tower_2d.local_attention.kv.1.weight")
    assert moved("vlm.vision_tower_3d.patch_embed.EncP.LGA_list.0.linear2.0.net1.0.weight")
    assert moved("vlm.vision_tower_3d.patch_embed.EncP.raw_point_embed.net.1.running_mean")
    assert moved("vlm.vision_tower_3d.proj.weight")
    assert not moved("vlm.vision_tower_2d.global_attention.proj.weight")
    assert not moved("vlm.vision_tower_3d.cls_token")
    assert torch.isfinite(trainer.grad_norm()).item()
