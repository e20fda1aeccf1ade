"""MoCo-v3-structure / v2-loss pretraining under SyncBatchNorm + DistributedDataParallel over NCCL (BASELINE configs[3],
SURVEY rows a13, a14, a16): the set-up of MAIN_PRE:273-312, the loop body of MAIN_PRE:510-548, and the self-checks that
make a multi-GPU run evidence rather than a timing:

  * every rank's queue is bit-identical after the steps, and queue_ptr advanced by world x batch per step (BLD:91-105,
    229-240: concat_all_gather of the keys is rank-major and lossless);
  * the loss is finite and the same on every rank's own shard statistics (SyncBN: global batch statistics);
  * step 1 of the data-parallel run reproduces ONE process running the concatenated global batch: this rank's rows of
    the logits, and the (DDP-averaged) gradients, against a single-process replica of the same model.

Used by bench.py (key "moco_dp", every N) and by tests/test_gpu_parity.py (world size 1 on the driver's single GPU).
"""
import copy
import importlib
import math
from functools import partial
from types import SimpleNamespace

import torch
import torch.distributed as dist
import torch.nn.functional as F

import e2e_common as E


def build_moco(seed=0, T=0.2):
    import vits
    bm = importlib.import_module("moco.builder_vit_mocov3structure_mocov2loss")
    torch.manual_seed(seed)  # identical replicas on every rank (DDP would broadcast rank 0's anyway)
    args = SimpleNamespace(arch="vit_small")
    return bm.MoCo_ViT(partial(vits.vit_small, stop_grad_conv1=True), args, 256, 4096, T)  # MAIN_PRE:273-275


def structured_views(B, hw, rank, device):
    """Two augmented 'views' per sample with real between-sample variation (smooth per-sample patterns + noise): iid
    noise images give nearly identical CLS tokens, which makes train-mode BatchNorm divide by ~0 and amplify rounding."""
    g = torch.Generator().manual_seed(4242 + rank)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, hw), torch.linspace(-1, 1, hw), indexing="ij")
    coef = torch.randn(B, 3, 6, generator=g)
    basis = torch.stack([torch.ones_like(xx), xx, yy, xx * yy, torch.sin(3 * xx), torch.cos(3 * yy)])  # [6,H,W]
    base = torch.einsum("bck,khw->bchw", coef, basis)
    q = base + 0.3 * torch.randn(B, 3, hw, hw, generator=g)
    k = base.flip(-1) + 0.3 * torch.randn(B, 3, hw, hw, generator=g)
    return q.to(device), k.to(device)


def _grad_cos(named_a, named_b):
    """min per-tensor gradient cosine; tensors whose gradient is zero up to rounding are left out (the final LayerNorm
    bias feeds a bias-free Linear + train-mode BatchNorm: its true gradient is exactly 0, what is there is 1e-8 noise)."""
    b = dict(named_b)
    named_a = list(named_a)
    top = max([float(p.grad.abs().max()) for _, p in named_a if p.grad is not None] + [0.0])
    worst, name = 1.0, None
    for n, p in named_a:
        if p.grad is None or float(p.grad.abs().max()) < 1e-6 * top:
            continue
        c = E.cos(p.grad, b[n].grad)
        if c < worst:
            worst, name = c, n
    return worst, name


def run(device, rank, world, batch=128, steps=10, warmup=3, check_batch=None, hw=224):
    """Returns a dict (rank 0's view; all ranks must call).  `check_batch` (default min(batch, 32)) is the per-rank batch
    of the step-1 parity check against a single-process run of the concatenated batch."""
    from mfvit.pretrain import MoCoPretrainer
    out = {"world": world, "batch_per_gpu": batch, "K": 65536, "steps": steps,
           "wrap": "SyncBatchNorm.convert_sync_batchnorm + DistributedDataParallel (NCCL), MAIN_PRE:297,312"}
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("moco_dp needs an initialised process group (MAIN_PRE only supports DDP, :318-323)")

    # ---------------------------------------------------------------- step-1 parity vs one process on the global batch
    cb = check_batch or min(batch, 32)
    model = build_moco().to(device)
    single = copy.deepcopy(model)  # same weights, BatchNorm1d, no DDP: fed the concatenated batch below
    model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)                                   # MAIN_PRE:297
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[device.index])              # MAIN_PRE:312
    im_q, im_k = structured_views(cb, hw, rank, device)
    m = 0.99
    logits, labels = ddp(im_q, im_k, m)
    F.cross_entropy(logits, labels).backward()
    gq = [torch.empty_like(im_q) for _ in range(world)]
    gk = [torch.empty_like(im_k) for _ in range(world)]
    dist.all_gather(gq, im_q)
    dist.all_gather(gk, im_k)
    single.train()
    # concat_all_gather inside a world-size-1 view of the same model: run it outside the process group's collectives
    single_logits, single_labels = _single_process_forward(single, torch.cat(gq), torch.cat(gk), m)
    F.cross_entropy(single_logits, single_labels).backward()
    mine = single_logits[rank * cb:(rank + 1) * cb]
    # the keys this step enqueued are part of neither logits (enqueue happens after the logits), so columns agree 1:1
    out["parity_logits_max_abs"] = float((logits.detach() - mine.detach()).abs().max())
    worst, name = _grad_cos(single.base_encoder.named_parameters(), ddp.module.base_encoder.named_parameters())
    w2, n2 = _grad_cos(single.predictor.named_parameters(), ddp.module.predictor.named_parameters())
    out["parity_grad_cos_min"] = min(worst, w2)
    out["parity_grad_worst"] = name if worst <= w2 else "predictor." + str(n2)
    # keys: per-row encoder results are bit-identical; SyncBatchNorm combines per-rank statistics in another order than
    # BatchNorm1d over the whole batch, so the normalised keys agree to rounding, not bit for bit
    out["parity_queue_max_abs"] = float((single.queue - ddp.module.queue).abs().max())
    del single, single_logits

    # ---------------------------------------------------------------- timed loop: MAIN_PRE:510-548 through MoCoPretrainer
    ddp.zero_grad(set_to_none=True)
    pre = MoCoPretrainer(ddp, lr=1.5e-4 * batch * world / 4 / 64, weight_decay=0.1, epochs=100, warmup_epochs=10, moco_m=0.99)
    im_q, im_k = structured_views(batch, hw, rank, device)
    iters_per_epoch = 100
    ptr0 = int(ddp.module.queue_ptr)
    for i in range(warmup):
        loss = pre.step(im_q, im_k, i / iters_per_epoch)
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        loss = pre.step(im_q, im_k, (warmup + i) / iters_per_epoch)
    b.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t) / steps
    out["ms_per_step"] = ms
    out["img_per_s"] = world * batch / ms * 1e3
    out["tflops_algorithmic"] = 4.74e12 * batch / 128 * world / ms / 1e9   # 4.74 TF per 128-image GPU step
    out["loss"] = float(loss)
    out["loss_finite"] = bool(math.isfinite(out["loss"]))
    # ---------------------------------------------------------------- self-checks on the queue
    q = ddp.module.queue
    sig = torch.stack([q.double().sum(), (q.double() * torch.arange(q.shape[1], device=device, dtype=torch.float64)).sum(),
                       q[:, :world * batch].double().abs().sum()])
    sigs = [torch.empty_like(sig) for _ in range(world)]
    dist.all_gather(sigs, sig)
    out["queues_identical_on_all_ranks"] = all(bool(torch.equal(s, sigs[0])) for s in sigs)
    expect = (ptr0 + (warmup + steps) * world * batch) % 65536
    out["queue_ptr"] = int(ddp.module.queue_ptr)
    out["queue_ptr_expected"] = expect
    norms = q[:, ptr0:ptr0 + world * batch].norm(dim=0)
    out["enqueued_keys_unit_norm"] = bool(((norms - 1).abs() < 1e-4).all())
    out["optimizer"] = "fused AdamW on flat buffers (mfv_adam_step_dev), bf16 autocast heads, no GradScaler"
    out["sync_batchnorm"] = ("%d nn.SyncBatchNorm modules replaced by mfvit.syncbn.FastSyncBatchNorm (same tensors; one "
                             "collective per call)" % pre.swapped_syncbn) if pre.swapped_syncbn else "torch.nn.SyncBatchNorm"
    out["ok"] = bool(out["queues_identical_on_all_ranks"] and out["queue_ptr"] == expect and out["loss_finite"]
                     and out["enqueued_keys_unit_norm"] and out["parity_queue_max_abs"] <= 1e-4
                     and out["parity_logits_max_abs"] <= 2e-3 and out["parity_grad_cos_min"] >= 0.999)
    del pre, ddp, model
    torch.cuda.empty_cache()
    return out


def _single_process_forward(model, im_q, im_k, m):
    """MoCo.forward as a 1-rank job would run it (no collectives): temporarily hide the process group from the builder
    module so concat_all_gather is the identity, as it is for world size 1."""
    bm = importlib.import_module("moco.builder_vit_mocov3structure_mocov2loss")
    saved = bm._dist_on
    bm._dist_on = lambda: False
    try:
        return model(im_q, im_k, m)
    finally:
        bm._dist_on = saved
