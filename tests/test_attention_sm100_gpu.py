"""tcgen05 attention forward (head_dim 128) vs the oracle and vs the mma.sync kernel it replaces."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,S,H,masked", [(1, 128, 1, False), (2, 150, 2, False), (2, 548, 3, False),
                                          (3, 548, 2, True), (1, 700, 2, True), (2, 64, 1, False), (1, 257, 1, True)])
def test_attn_fwd_sm100(cuda_lib, B, S, H, masked):
    from mla_b200 import ops
    from oracle import llama as O
    torch.manual_seed(11)
    D = 128
    h = H * D
    qkv = torch.randn(B * S, 3 * h, device="cuda").to(torch.bfloat16)
    mask = None
    if masked:
        mask = torch.ones(B, S, dtype=torch.bool, device="cuda")
        mask[0, S - 9:] = False
        if B > 1:
            mask[1, S // 3: S // 3 + 5] = False
    ops.ATTN_IMPL["fwd"] = "mma"
    ctx_ref, lse_ref = ops.attn_fwd(qkv, B, S, H, D, mask)
    ops.ATTN_IMPL["fwd"] = "sm100"
    ctx, lse = ops.attn_fwd(qkv, B, S, H, D, mask)
    torch.cuda.synchronize()
    q, k, v = [qkv[:, i * h:(i + 1) * h].float().reshape(B, S, H, D).transpose(1, 2) for i in range(3)]
    truth = O.attention(q, k, v, mask).reshape(B * S, h)
    e_new, e_old = rel_err(ctx, truth), rel_err(ctx_ref, truth)
    assert e_new < 1.3 * e_old + 1e-3, (e_new, e_old)
    assert rel_err(ctx, ctx_ref) < 6e-3
    fin = torch.isfinite(lse_ref)
    assert torch.equal(torch.isfinite(lse), fin)
    assert (lse[fin] - lse_ref[fin]).abs().max() < 2e-3
    if masked:
        assert ctx.view(B, S, h)[0, S - 9:].abs().max() == 0


def test_attn_fwd_sm100_speed(cuda_lib):
    """Not a pass/fail on speed: prints both kernels' time at the benchmark shape for the log."""
    from mla_b200 import ops
    B, S, H, D = 32, 548, 32, 128
    qkv = torch.randn(B * S, 3 * H * D, device="cuda").to(torch.bfloat16)
    res = {}
    for impl in ("mma", "sm100"):
        ops.ATTN_IMPL["fwd"] = impl
        for _ in range(3):
            ops.attn_fwd(qkv, B, S, H, D, None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.attn_fwd(qkv, B, S, H, D, None)
        e1.record()
        torch.cuda.synchronize()
        res[impl] = e0.elapsed_time(e1) / 10
    ops.ATTN_IMPL["fwd"] = "sm100"
    flops = 4.0 * S * S * D * H * B / 2
    print(f"\nattn fwd [32,548,32,128]: mma.sync {res['mma']:.3f} ms ({flops / res['mma'] / 1e9:.0f} TFLOP/s causal), "
          f"tcgen05 {res['sm100']:.3f} ms ({flops / res['sm100'] / 1e9:.0f} TFLOP/s)")


@pytest.mark.parametrize("B,S,H,masked", [(1, 128, 1, False), (2, 150, 2, False), (2, 548, 2, False),
                                          (2, 548, 2, True), (1, 700, 1, True), (1, 64, 1, False), (1, 257, 2, True)])
def test_attn_bwd_sm100(cuda_lib, B, S, H, masked):
    from mla_b200 import ops
    from oracle import llama as O
    torch.manual_seed(12)
    D = 128
    h = H * D
    qkv = torch.randn(B * S, 3 * h, device="cuda").to(torch.bfloat16)
    mask = None
    if masked:
        mask = torch.ones(B, S, dtype=torch.bool, device="cuda")
        mask[0, S - 9:] = False
        if B > 1:
            mask[1, S // 3: S // 3 + 5] = False
    ctx, lse = ops.attn_fwd(qkv, B, S, H, D, mask)
    dctx = torch.randn(B * S, h, device="cuda").to(torch.bfloat16)
    keep = dict(ops.ATTN_IMPL)
    ops.ATTN_IMPL["bwd"] = "mma"
    d_ref = ops.attn_bwd(dctx, qkv, ctx, lse, B, S, H, D, mask)
    ops.ATTN_IMPL["bwd"] = "sm100"
    d_new = ops.attn_bwd(dctx, qkv, ctx, lse, B, S, H, D, mask)
    ops.ATTN_IMPL.update(keep)
    torch.cuda.synchronize()
    qf = qkv.float().requires_grad_(True)
    q, k, v = [qf[:, i * h:(i + 1) * h].reshape(B, S, H, D).transpose(1, 2) for i in range(3)]
    O.attention(q, k, v, mask).backward(dctx.view(B, S, h).float())
    for i, name in enumerate("qkv"):
        sl = slice(i * h, (i + 1) * h)
        e_new, e_old = rel_err(d_new[:, sl], qf.grad[:, sl]), rel_err(d_ref[:, sl], qf.grad[:, sl])
        assert e_new < 1.3 * e_old + 2e-3, (name, e_new, e_old)
    assert torch.isfinite(d_new.float()).all()


@pytest.mark.parametrize("B,S,H,masked", [(1, 128, 1, False), (2, 150, 2, False), (2, 548, 2, False), (2, 548, 2, True),
                                          (1, 700, 1, True), (1, 64, 1, False), (1, 257, 2, True), (3, 1060, 2, False),
                                          (1, 3876, 1, False), (1, 40, 1, False)])
def test_attn_bwd_pipelined_matches_first_generation(cuda_lib, B, S, H, masked):
    """attention_bwd2_sm100.cu (two math groups, 3-slot ring, double-buffered staging) computes the same products in the
    same order as attention_bwd_sm100.cu: bit-identical dq | dk | dv; with the RoPE transpose fused into its epilogue it
    equals the first generation followed by the in-place transposed rotation, bit for bit."""
    from mla_b200 import ops
    torch.manual_seed(13 + S)
    D = 128
    h = H * D
    qkv = torch.randn(B * S, 3 * h, device="cuda").to(torch.bfloat16)
    mask = None
    if masked:
        mask = torch.ones(B, S, dtype=torch.bool, device="cuda")
        mask[0, S - 9:] = False
        if B > 1:
            mask[1, S // 3: S // 3 + 5] = False
    ctx, lse = ops.attn_fwd(qkv, B, S, H, D, mask)
    dctx = torch.randn(B * S, h, device="cuda").to(torch.bfloat16)
    ang = torch.rand(S, D // 2, device="cuda") * 6.28
    cos, sin = ang.cos().bfloat16().contiguous(), ang.sin().bfloat16().contiguous()
    keep = dict(ops.ATTN_IMPL)
    try:
        ops.ATTN_IMPL["bwd"] = "sm100"
        want = ops.attn_bwd(dctx, qkv, ctx, lse, B, S, H, D, mask)
        ops.ATTN_IMPL["bwd"] = "sm100v2"
        want_r = want.clone()
        ops.rope_(want_r, 0, 2 * H, D, S, cos, sin, transpose=True)
        for ts in (0, 1):       # P/dS staged in shared memory | handed over in tensor memory (A operand from TMEM)
            cuda_lib.mla_attn_bwd2_set_ts(ts)
            for _ in range(3):      # several launches back to back: barrier phases / TMEM re-allocation
                got = ops.attn_bwd(dctx, qkv, ctx, lse, B, S, H, D, mask)
            torch.cuda.synchronize()
            assert torch.equal(got, want), ts
            got_r = ops.attn_bwd(dctx, qkv, ctx, lse, B, S, H, D, mask, rope=(cos, sin))
            torch.cuda.synchronize()
            assert torch.equal(got_r, want_r), ts
    finally:
        cuda_lib.mla_attn_bwd2_set_ts(1)
        ops.ATTN_IMPL.update(keep)


def test_attn_bwd_sm100_speed(cuda_lib):
    from mla_b200 import ops
    B, S, H, D = 32, 548, 32, 128
    qkv = torch.randn(B * S, 3 * H * D, device="cuda").to(torch.bfloat16)
    ctx, lse = ops.attn_fwd(qkv, B, S, H, D, None)
    dctx = torch.randn_like(ctx)
    res = {}
    keep = dict(ops.ATTN_IMPL)
    for impl in ("mma", "sm100", "sm100v2s", "sm100v2"):
        ops.ATTN_IMPL["bwd"] = impl.rstrip("s") if impl.startswith("sm100v2") else impl
        cuda_lib.mla_attn_bwd2_set_ts(0 if impl == "sm100v2s" else 1)
        for _ in range(3):
            ops.attn_bwd(dctx, qkv, ctx, lse, B, S, H, D, None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.attn_bwd(dctx, qkv, ctx, lse, B, S, H, D, None)
        e1.record()
        torch.cuda.synchronize()
        res[impl] = e0.elapsed_time(e1) / 10
    ops.ATTN_IMPL.update(keep)
    fl = 2.5 * 4.0 * S * S * D * H * B / 2
    print(f"\nattn bwd [32,548,32,128]: mma.sync {res['mma']:.3f} ms, tcgen05 gen 1 {res['sm100']:.3f} ms "
          f"({fl / res['sm100'] / 1e9:.0f} TFLOP/s), pipelined/smem {res['sm100v2s']:.3f} ms "
          f"({fl / res['sm100v2s'] / 1e9:.0f} TFLOP/s), pipelined/TMEM {res['sm100v2']:.3f} ms ({fl / res['sm100v2'] / 1e9:.0f} TFLOP/s)")
