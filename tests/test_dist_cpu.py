"""CPU, world_size=2 over gloo: host-side logic of the data-parallel path - rank-major key gathering,
queue pointer advance, gradient averaging on the flat buffers, per-rank data sharding.  No CUDA involved."""
import importlib
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "multi-feature-vit_b200"), os.path.join(ROOT, "multi-feature-vit_b200", "dropin"),
              os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import moco_ref
    bm = importlib.import_module("moco.builder_vit_mocov3structure_mocov2loss")
    res = {}
    # (1) rank-major concatenation: oracle restatement == drop-in implementation (BLD:229-240)
    x = torch.full((2, 3), float(rank)) + torch.arange(3.0)
    a = moco_ref.concat_all_gather(x)
    b = bm.concat_all_gather(x)
    res["gather_equal"] = bool(torch.equal(a, b))
    res["gather_order"] = a[:, 0].tolist()
    # (2) enqueue on every rank with gathered keys -> identical queues, ptr += world * batch (BLD:91-105)
    K, D, B = 16, 4, 2
    queue, ptr = torch.zeros(D, K), torch.zeros(1, dtype=torch.long)
    keys = torch.randn(B, D, generator=torch.Generator().manual_seed(100 + rank))
    moco_ref.dequeue_and_enqueue(queue, ptr, keys, K)
    res["ptr"] = int(ptr)
    gathered = [torch.zeros_like(queue) for _ in range(world)]
    dist.all_gather(gathered, queue)
    res["queues_identical"] = bool(all(torch.equal(gathered[0], g) for g in gathered))
    res["queue_cols"] = queue[:, :world * B].t().tolist()
    res["own_keys"] = keys.tolist()
    # (3) gradient averaging of the trainer on flat buffers (mean over ranks, DDP semantics)
    from mfvit.trainer import MFViTCATrainer

    class _Small:
        pass
    tr = MFViTCATrainer.__new__(MFViTCATrainer)
    tr.pg = None
    tr._small = _Small()
    tr._small.grad = torch.full((8,), float(rank + 1))
    big = torch.full((2, 16), float(10 * (rank + 1)))
    tr.all_reduce(big)
    res["grad_mean_big"] = float(big[0, 0])
    res["grad_mean_small"] = float(tr._small.grad[0])
    tr.local_only = True  # bench.py's same-work reference: no collective, gradients stay local
    local = torch.full((2, 4), float(rank + 1))
    tr.all_reduce(local)
    res["local_only_untouched"] = bool(torch.equal(local, torch.full((2, 4), float(rank + 1))))
    # (4) per-rank synthetic shards differ, same rank reproduces (SURVEY 8(d) seeding)
    import e2e_common as E
    c0, _, t0 = E.synthetic_pair(2, 32, rank=rank)
    c1, _, _ = E.synthetic_pair(2, 32, rank=rank)
    res["shard_reproducible"] = bool(torch.equal(c0, c1))
    res["shard_sum"] = float(c0.sum())
    # (5) FastSyncBatchNorm (mfvit/syncbn.py): two ranks with UNEQUAL batches reproduce nn.BatchNorm1d on the concatenated
    # batch - output, running statistics, input gradient; parameter gradients are local sums that add up to the global one
    from mfvit.syncbn import FastSyncBatchNorm, swap_sync_batchnorm
    import torch.nn as nn
    Fdim = 6
    gen = torch.Generator().manual_seed(7)
    full = torch.randn(7, Fdim, generator=gen) * 3 + 50.0            # large mean: E[x^2] - mean^2 would lose digits
    wts = torch.randn(7, Fdim, generator=gen)
    lo, hi = (0, 3) if rank == 0 else (3, 7)
    ref_bn = nn.BatchNorm1d(Fdim)
    with torch.no_grad():
        ref_bn.weight.copy_(torch.linspace(0.5, 1.5, Fdim)); ref_bn.bias.copy_(torch.linspace(-1, 1, Fdim))
    stock = nn.SyncBatchNorm(Fdim)
    stock.load_state_dict(ref_bn.state_dict())
    holder = nn.Sequential(stock)
    res["bn_swapped"] = swap_sync_batchnorm(holder)
    fast = holder[0]
    res["bn_shares_tensors"] = bool(fast.weight is stock.weight and fast.running_mean is stock.running_mean)
    xr = full.clone().requires_grad_(True)
    (ref_bn(xr) * wts).sum().backward()
    xl = full[lo:hi].clone().requires_grad_(True)
    yl = fast(xl)
    (yl * wts[lo:hi]).sum().backward()
    want = nn.functional.batch_norm(full, None, None, ref_bn.weight, ref_bn.bias, True, 0.0, ref_bn.eps)  # no side effects
    res["bn_out_err"] = float((yl.detach() - want[lo:hi]).abs().max())
    res["bn_dx_err"] = float((xl.grad - xr.grad[lo:hi]).abs().max())
    gw = fast.weight.grad.clone()
    dist.all_reduce(gw)
    res["bn_dw_err"] = float((gw - ref_bn.weight.grad).abs().max())
    res["bn_rm_err"] = float((fast.running_mean - ref_bn.running_mean).abs().max())
    res["bn_rv_err"] = float((fast.running_var - ref_bn.running_var).abs().max())
    res["bn_tracked"] = int(fast.num_batches_tracked)
    res["bn_keys"] = sorted(fast.state_dict().keys())
    fast.eval()
    res["bn_eval_err"] = float((fast(full[lo:hi]) - ref_bn.eval()(full)[lo:hi]).abs().max())
    torch.save(res, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    world, port = 2, 29541 + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(str(tmp_path), "rank%d.pt" % i)) for i in range(world)]
    for x in r:
        assert x["gather_equal"]
        assert x["gather_order"] == [0.0, 0.0, 1.0, 1.0]          # rank-major
        assert x["ptr"] == 4 and x["queues_identical"]
        assert x["grad_mean_big"] == pytest.approx(15.0) and x["grad_mean_small"] == pytest.approx(1.5)
        assert x["local_only_untouched"]
        assert x["shard_reproducible"]
        assert x["bn_swapped"] == 1 and x["bn_shares_tensors"] and x["bn_tracked"] == 1
        assert x["bn_keys"] == ["bias", "num_batches_tracked", "running_mean", "running_var", "weight"]
        assert x["bn_out_err"] < 2e-5 and x["bn_dx_err"] < 2e-5 and x["bn_dw_err"] < 1e-4
        assert x["bn_rm_err"] < 1e-5 and x["bn_rv_err"] < 1e-5 and x["bn_eval_err"] < 2e-5
    # the queue holds rank 0's keys first, then rank 1's
    assert torch.allclose(torch.tensor(r[0]["queue_cols"][:2]), torch.tensor(r[0]["own_keys"]))
    assert torch.allclose(torch.tensor(r[0]["queue_cols"][2:]), torch.tensor(r[1]["own_keys"]))
    assert r[0]["shard_sum"] != r[1]["shard_sum"]
