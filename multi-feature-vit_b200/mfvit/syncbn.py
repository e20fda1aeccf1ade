"""Lean synchronised BatchNorm for the MoCo projector / predictor MLPs under data parallelism.

MAIN_PRE:297 converts every BatchNorm of the model with torch.nn.SyncBatchNorm.convert_sync_batchnorm.  The stock module
costs about 1 ms of host time per call (batch_norm_stats, an all_gather work object, batch_norm_gather_stats_with_counts,
the element-wise pass and an autograd node, each dispatched separately): with ten forward and five backward calls per
MoCo step the step becomes host-bound (measured on two B200s, 128 images per GPU and view: 12.7 ms without any wrapper,
14.7 ms with DDP, 17.1 ms with SyncBatchNorm, 19.3 ms with both).  These activations are tiny ([128, 4096] and [128, 256]),
so the whole layer is a handful of fused element-wise launches around ONE all_gather (forward) and ONE all_reduce
(backward) of a [3, F] / [2, F] fp32 tensor.

`swap_sync_batchnorm(model)` replaces each nn.SyncBatchNorm by a FastSyncBatchNorm that SHARES its parameters and buffers
(state-dict keys, DDP's parameter hooks and buffer broadcasts are untouched).  MoCoPretrainer does this by default; the
reference's own loop over the drop-in keeps the stock module.

Statistics: every rank sends (mean_i, M2_i, n_i) of its rows; the global mean and biased variance are combined with
Chan's formula (no E[x^2] - mean^2 cancellation), exactly what batch_norm_gather_stats_with_counts computes.  Running
statistics use the unbiased variance, as nn.BatchNorm does.  Parameter gradients are local sums (DDP averages them, as
with the stock module); the input gradient uses the global sums.
"""
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F


def _world(group):
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    return dist.get_world_size(group)


def _all_gather_rows(t, group):
    """[..] -> [world, ..], rank-major."""
    world = _world(group)
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    if t.is_cuda:
        dist.all_gather_into_tensor(out.view(world * t.shape[0], *t.shape[1:]), t.contiguous(), group=group)
    else:  # gloo (CPU tests): the list form works on every backend
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous(), group=group)
        out.copy_(torch.stack(parts))
    return out


class _SyncBNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, group):
        xf = x.float()
        n_local = xf.shape[0]
        mean_l = xf.mean(0)
        m2_l = (xf - mean_l).square().sum(0)
        stats = torch.stack([mean_l, m2_l, torch.full_like(mean_l, float(n_local))])           # [3, F]
        allst = _all_gather_rows(stats, group)                                                  # [world, 3, F]
        cnt = allst[:, 2]
        n = cnt.sum(0)                                                                          # [F], all equal
        mean = (allst[:, 0] * cnt).sum(0) / n
        m2 = allst[:, 1].sum(0) + (cnt * (allst[:, 0] - mean).square()).sum(0)
        var = m2 / n
        rstd = torch.rsqrt(var + eps)
        xhat = (xf - mean) * rstd
        y = xhat if weight is None else torch.addcmul(bias, xhat, weight) if bias is not None else xhat * weight
        if running_mean is not None:
            with torch.no_grad():
                running_mean.mul_(1 - momentum).add_(mean.to(running_mean.dtype), alpha=momentum)
                running_var.mul_(1 - momentum).add_((m2 / (n - 1).clamp_min(1)).to(running_var.dtype), alpha=momentum)
        ctx.save_for_backward(xhat, rstd, weight, n)
        ctx.group = group
        ctx.in_dtype = x.dtype
        ctx.has_bias = bias is not None
        return y.to(x.dtype)

    @staticmethod
    def backward(ctx, dy):
        xhat, rstd, weight, n = ctx.saved_tensors
        dyf = dy.float()
        d_weight = (dyf * xhat).sum(0) if weight is not None and ctx.needs_input_grad[1] else None
        d_bias = dyf.sum(0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        dx = None
        if ctx.needs_input_grad[0]:
            g = dyf if weight is None else dyf * weight
            sums = torch.stack([g.sum(0), (g * xhat).sum(0)])                                   # [2, F] local
            if _world(ctx.group) > 1:
                dist.all_reduce(sums, group=ctx.group)
            dx = ((g - sums[0] / n - xhat * (sums[1] / n)) * rstd).to(ctx.in_dtype)
        return dx, d_weight, d_bias, None, None, None, None, None


class FastSyncBatchNorm(nn.Module):
    """Drop-in for nn.SyncBatchNorm on [N, F] inputs (the MLP heads); same parameters, buffers and state-dict keys."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True, process_group=None):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.affine, self.track_running_stats, self.process_group = affine, track_running_stats, process_group
        if affine:
            self.weight = nn.Parameter(torch.ones(num_features))
            self.bias = nn.Parameter(torch.zeros(num_features))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)
        if track_running_stats:
            self.register_buffer("running_mean", torch.zeros(num_features))
            self.register_buffer("running_var", torch.ones(num_features))
            self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        else:
            self.register_buffer("running_mean", None)
            self.register_buffer("running_var", None)
            self.register_buffer("num_batches_tracked", None)

    @classmethod
    def from_module(cls, bn):
        """Adopt the parameters and buffers of an nn.SyncBatchNorm / nn.BatchNorm1d (the same tensor objects)."""
        new = cls.__new__(cls)
        nn.Module.__init__(new)
        new.num_features, new.eps, new.momentum = bn.num_features, bn.eps, bn.momentum
        new.affine, new.track_running_stats = bn.affine, bn.track_running_stats
        new.process_group = getattr(bn, "process_group", None)
        new.register_parameter("weight", bn.weight)
        new.register_parameter("bias", bn.bias)
        new.register_buffer("running_mean", bn.running_mean)
        new.register_buffer("running_var", bn.running_var)
        new.register_buffer("num_batches_tracked", bn.num_batches_tracked)
        new.train(bn.training)
        return new

    def forward(self, x):
        if x.dim() != 2:
            raise ValueError("FastSyncBatchNorm handles [N, F] inputs (MLP heads); got %s" % (tuple(x.shape),))
        use_batch_stats = self.training or self.running_mean is None
        if not use_batch_stats or _world(self.process_group) == 1:
            if self.training and self.num_batches_tracked is not None:
                self.num_batches_tracked.add_(1)
            return F.batch_norm(x, self.running_mean, self.running_var, self.weight, self.bias, use_batch_stats,
                                self.momentum if self.momentum is not None else 0.0, self.eps)
        if self.momentum is None:
            raise ValueError("cumulative moving average (momentum=None) is not supported")
        if self.num_batches_tracked is not None:
            self.num_batches_tracked.add_(1)
        rm = self.running_mean if self.training else None
        rv = self.running_var if self.training else None
        return _SyncBNFn.apply(x, self.weight, self.bias, rm, rv, self.eps, self.momentum, self.process_group)

    def extra_repr(self):
        return "{num_features}, eps={eps}, momentum={momentum}, affine={affine}".format(**self.__dict__)


def swap_sync_batchnorm(model):
    """Replace every nn.SyncBatchNorm below `model` (a DDP wrapper is fine) by a FastSyncBatchNorm sharing its tensors.
    Returns the number of modules replaced."""
    n = 0
    for parent in list(model.modules()):
        for name, child in list(parent.named_children()):
            if isinstance(child, nn.SyncBatchNorm):
                setattr(parent, name, FastSyncBatchNorm.from_module(child))
                n += 1
    return n
