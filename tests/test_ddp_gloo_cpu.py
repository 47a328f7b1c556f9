"""world_size-2 gloo test (CPU) of the data-parallel host logic: the per-layer gradient arenas and the small modules'
gradients are all-reduced (sum) across ranks, parameters without gradients are skipped consistently, and per-rank
synthetic batches differ (seed = 1234 + rank) while the model weights are identical."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mla_b200.llama import LlamaModel
        from mla_b200.synthetic import make_batch
        from mla_b200.trainer import DataParallelTrainer
        torch.manual_seed(0)                                   # identical replicas
        model = LlamaModel(64, 32, 64, 2, 4)
        for p_ in model.parameters():
            torch.nn.init.normal_(p_, std=0.02)
        extra = torch.nn.Linear(4, 4)
        unused = torch.nn.Linear(4, 4)                         # never gets a gradient (like lm_head in diffusion mode)
        root = torch.nn.ModuleDict(dict(llm=model, extra=extra, unused=unused))
        tr = DataParallelTrainer(root, lr=1e-3)
        assert tr.world == world and all(l._grad_ready_cb is not None for l in tr.layers)
        # fake a backward: every layer fills its arenas with a rank-dependent value and signals readiness
        for li, layer in enumerate(tr.layers):
            layer.grad_arenas()
            layer._attach_grads()
            for g in layer._g:
                g.fill_(float(rank + 1) * (li + 1))
            layer._grads_fresh = False
            layer._grad_ready_cb(layer)
        extra.weight.grad = torch.full_like(extra.weight, float(rank + 1))
        extra.bias.grad = torch.full_like(extra.bias, 10.0 * (rank + 1))
        tr.exchange()
        tot = sum(r + 1 for r in range(world))
        for li, layer in enumerate(tr.layers):
            for g in layer._g:
                assert torch.all(g == tot * (li + 1)), (rank, li)
            assert layer.self_attn.q_proj.weight.grad is layer._views[0]      # param.grad aliases the arena
        assert torch.all(extra.weight.grad == tot) and torch.all(extra.bias.grad == 10.0 * tot)
        assert unused.weight.grad is None and model.embed_tokens.weight.grad is None
        assert not tr._handles
        # replicas see different data, same weights
        b = make_batch(2, 8, seed=1234 + rank, image_hw=42)
        gathered = [torch.zeros_like(b["actions"]) for _ in range(world)]
        dist.all_gather(gathered, b["actions"])
        assert not torch.equal(gathered[0], gathered[1])
        w = model.layers[0].mlp.down_proj.weight.detach().clone()
        ws = [torch.zeros_like(w) for _ in range(world)]
        dist.all_gather(ws, w)
        assert torch.equal(ws[0], ws[1])
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def _accum_worker(rank, world, port, out_dir):
    """Gradient accumulation (no_sync), the double-backward guard, the bf16 reduce option, the rank-0 broadcast."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mla_b200.llama import LlamaModel
        from mla_b200.trainer import DataParallelTrainer
        torch.manual_seed(100 + rank)                          # replicas built DIFFERENTLY: the trainer must align them
        model = LlamaModel(64, 32, 64, 2, 4)
        for p_ in model.parameters():
            torch.nn.init.normal_(p_, std=0.02)
        model.register_buffer("running_stat", torch.full((3,), float(rank)))
        extra = torch.nn.Linear(4, 4)
        root = torch.nn.ModuleDict(dict(llm=model, extra=extra))
        for reduce_dtype in (None, torch.bfloat16):
            tr = DataParallelTrainer(root, lr=1e-3, reduce_dtype=reduce_dtype)
            w = model.layers[1].mlp.up_proj.weight.detach().clone()
            ws = [torch.zeros_like(w) for _ in range(world)]
            dist.all_gather(ws, w)
            assert torch.equal(ws[0], ws[1]), "parameters were not broadcast from rank 0"
            assert torch.all(model.running_stat == 0.0), "buffers were not broadcast from rank 0"

            def fake_backward(val):
                for li, layer in enumerate(tr.layers):
                    layer.grad_arenas()
                    live = layer._attach_grads()
                    for g in layer._g:
                        if live:
                            g.add_(val * (li + 1))
                        else:
                            g.fill_(val * (li + 1))
                    layer._grads_fresh = False
                    layer._grad_ready_cb(layer)

            with tr.no_sync():
                fake_backward(float(rank + 1))                 # micro-batch 1: local only
            assert not tr._handles and not tr._reduced
            fake_backward(10.0 * (rank + 1))                   # micro-batch 2: reduces the accumulated sums
            assert len(tr._handles) == len(tr.layers)          # ONE collective per decoder layer
            extra.weight.grad = torch.full_like(extra.weight, float(rank + 1))
            tr.exchange()
            tot = sum(11.0 * (r + 1) for r in range(world))
            for li, layer in enumerate(tr.layers):
                for g in layer._g:
                    assert torch.all(g == tot * (li + 1)), (rank, li, g.flatten()[0].item(), tot * (li + 1))
            assert torch.all(extra.weight.grad == sum(r + 1 for r in range(world)))
            # a further backward in the same step would add onto rank sums: refused loudly
            with pytest.raises(RuntimeError, match="no_sync"):
                fake_backward(1.0)
            tr._reduced.clear()
            for layer in tr.layers:
                layer.mark_grads_fresh()
            with pytest.raises(RuntimeError, match="no_sync"):
                with tr.no_sync():
                    tr.exchange()
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_gradient_accumulation_and_broadcast_world2(tmp_path):
    world = 2
    mp.spawn(_accum_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_lr_schedule_matches_reference():
    """linear-warmup+cosine-decay == transformers.get_cosine_schedule_with_warmup as the reference steps it
    (fsdp.py:258-260: lr starts at 0; the k-th optimizer step runs at lambda(k-1)); other names raise like the reference."""
    import math
    from mla_b200.llama import LlamaModel
    from mla_b200.trainer import DataParallelTrainer
    m = LlamaModel(64, 32, 64, 1, 4)
    with pytest.raises(ValueError, match="not supported"):
        DataParallelTrainer(m, lr_scheduler_type="cosine")
    with pytest.raises(ValueError, match="total_steps"):
        DataParallelTrainer(m, lr_scheduler_type="linear-warmup+cosine-decay", warmup_steps=3)
    tr = DataParallelTrainer(m, lr=2e-5, lr_scheduler_type="linear-warmup+cosine-decay", warmup_steps=3, total_steps=20)
    ref = torch.optim.AdamW([torch.nn.Parameter(torch.zeros(1))], lr=2e-5)
    sched = torch.optim.lr_scheduler.LambdaLR(ref, lambda s: (s / 3.0) if s < 3 else max(
        0.0, 0.5 * (1.0 + math.cos(math.pi * (s - 3) / 17.0))))       # the lambda of get_cosine_schedule_with_warmup
    for k in range(1, 21):
        tr.step_count = k
        assert abs(tr.current_lr() - ref.param_groups[0]["lr"]) < 1e-12, k
        ref.step()
        sched.step()
    assert DataParallelTrainer(m, lr=1e-4).current_lr() == 1e-4


def test_gradient_exchange_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def _ckpt_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mla_b200.llama import LlamaModel
        from mla_b200.trainer import DataParallelTrainer
        torch.manual_seed(0)

        class Wrap(torch.nn.Module):          # the attributes save_checkpoint reads from the MLA wrapper
            def __init__(self):
                super().__init__()
                self.vlm = torch.nn.ModuleDict(dict(llm_backbone=LlamaModel(64, 32, 64, 2, 4),
                                                    projector_2d=torch.nn.Linear(4, 4)))
                self.trainable_module_keys = ["vlm.llm_backbone", "vlm.projector_2d"]
                self.all_module_keys = list(self.trainable_module_keys)
        m = Wrap()
        tr = DataParallelTrainer(m)
        tr.step_count = 3 + rank                          # only rank 0's optimizer record is written
        path = tr.save_checkpoint(out_dir, global_step=5, epoch=1, train_loss=0.5)
        assert path.name == "step-000005-epoch-01-loss=0.5000.pt"
        assert path.exists()                               # visible to every rank after the barrier
        blob = torch.load(path, weights_only=True)["model"]
        assert set(blob) == {"llm_backbone", "projector_2d"}
        opt = torch.load(path.with_suffix(".optimizer"), weights_only=True)
        assert opt["optimizer"]["step"] == 3 and opt["scheduler"] == {"epoch": 1, "global_step": 5}
        open(os.path.join(out_dir, f"ckpt_ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_checkpoint_written_once_world2(tmp_path):
    """Rank 0 writes its replica (no gather), every rank leaves save_checkpoint behind the same barrier."""
    world = 2
    mp.spawn(_ckpt_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ckpt_ok{r}").exists() for r in range(world))
    assert len(list((tmp_path / "checkpoints").glob("*.pt"))) == 1
