"""autograd seams over the fused kernels that sit outside the encoder engine: the small classification head, the
CLS cross-attention fusion, the InfoNCE logits."""
import torch

from . import ops
from ._lib import FusionGrads

FUSION_FIELDS = ["ln1_w", "ln1_b", "wq", "wk", "wv", "proj_w", "proj_b", "ln2_w", "ln2_b", "head_w", "head_b",
                 "vhead_w", "vhead_b"]


class HeadFn(torch.autograd.Function):
    """y = Linear(C, N<=32) applied to the CLS row of tok [B,S,C] (MAIN_CA:309-310 / MAIN_LPFT:288 heads)."""

    @staticmethod
    def forward(ctx, tok, weight, bias):
        B, S, Cd = tok.shape
        tok = tok.contiguous()
        y = ops.linear_small_fwd(tok, S * Cd, weight.detach().contiguous(), None if bias is None else bias.detach(), B)
        ctx.save_for_backward(tok, weight, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        tok, weight, bias = ctx.saved_tensors
        B, S, Cd = tok.shape
        dtok = torch.empty_like(tok)
        ops.fill_(dtok, 0.0)
        dw = torch.empty_like(weight)
        ops.fill_(dw, 0.0) if dw.numel() % 4 == 0 else dw.zero_()
        db = torch.zeros_like(bias) if bias is not None else None
        ops.linear_small_bwd(tok, S * Cd, weight.detach().contiguous(), dy.contiguous(), dtok, S * Cd, dw, db, B)
        return dtok, dw, db


class FusionFn(torch.autograd.Function):
    """(fused [B,NC], x [2,B,NC]) = fusion(tok [2,B,S,C]); 26 parameter tensors in FUSION_FIELDS x direction order."""

    @staticmethod
    def forward(ctx, tok, heads, *params):
        _, B, S, Cd = tok.shape
        tok = tok.contiguous()
        ps = [None if p is None else p.detach().contiguous() for p in params]
        NC = ps[FUSION_FIELDS.index("head_w") * 2].shape[0]
        struct = ops.fusion_param_struct({n: (ps[2 * i], ps[2 * i + 1]) for i, n in enumerate(FUSION_FIELDS)})
        scratch = ops.fusion_scratch(tok, B, S, Cd, heads)  # forward state for the backward + its own records
        fused, x = ops.fusion_fwd(tok, struct, B, S, Cd, heads, NC, saved=scratch)
        ctx.scratch = scratch
        ctx.save_for_backward(tok, *[p for p in ps if p is not None])
        ctx.mask = [p is not None for p in ps]
        ctx.heads, ctx.NC = heads, NC
        return fused, x

    @staticmethod
    def backward(ctx, d_fused, d_x):
        tok = ctx.saved_tensors[0]
        it = iter(ctx.saved_tensors[1:])
        ps = [next(it) if m else None for m in ctx.mask]
        _, B, S, Cd = tok.shape
        # one flat zeroed buffer for all parameter gradients
        sizes = [0 if p is None else (p.numel() + 3) // 4 * 4 for p in ps]
        flat = torch.empty(max(sum(sizes), 4), device=tok.device, dtype=torch.float32)
        ops.fill_(flat, 0.0)
        gs, off = [], 0
        for p, n in zip(ps, sizes):
            gs.append(None if p is None else flat[off:off + p.numel()].view(p.shape))
            off += n
        struct = ops.fusion_param_struct({n: (ps[2 * i], ps[2 * i + 1]) for i, n in enumerate(FUSION_FIELDS)})
        gstruct = ops.fusion_param_struct({n: (gs[2 * i], gs[2 * i + 1]) for i, n in enumerate(FUSION_FIELDS)},
                                          cls=FusionGrads)
        if d_fused is None:
            d_fused = torch.zeros(B, ctx.NC, device=tok.device)
        dxc = None if d_x is None else d_x.contiguous()
        dtok = ops.fusion_bwd(tok, struct, gstruct, d_fused.contiguous(), dxc, B, S, Cd, ctx.heads, ctx.NC,
                              scratch=ctx.scratch)
        return (dtok, None) + tuple(gs)


class InfoNCEFn(torch.autograd.Function):
    """logits [N,1+K] = cat(q.k, q @ queue)/T with q,k L2-normalised in-kernel (BLD:165,175,183-191).

    `holder` is a dict the caller fills with {"start": ptr, "old": queue[:, ptr:ptr+n].clone()} right before it
    enqueues new keys, so that the backward uses the queue contents the forward saw (the reference clones the whole
    64 MiB queue every step for this, BLD:185)."""

    @staticmethod
    def forward(ctx, q_raw, k_raw, queue, T, holder):
        q_raw = q_raw.float().contiguous()
        k_raw = k_raw.float().contiguous()
        qn, kn, logits, lse, loss = ops.infonce_fwd(q_raw, k_raw, queue, T)
        ctx.save_for_backward(q_raw, qn, kn, logits, lse)
        ctx.queue, ctx.T, ctx.holder = queue, T, holder
        ctx.mark_non_differentiable(kn)
        return logits, kn, loss

    @staticmethod
    def backward(ctx, dlogits, _dkn, dloss):
        q_raw, qn, kn, logits, lse = ctx.saved_tensors
        h = ctx.holder or {}
        ov, start = h.get("old"), h.get("start", 0)
        if dlogits is not None:
            dq = ops.infonce_bwd(q_raw, qn, kn, ctx.queue, logits, lse, ctx.T, dlogits=dlogits.contiguous(),
                                 override=ov, ov_start=start)
            if dloss is not None:
                dq = dq + ops.infonce_bwd(q_raw, qn, kn, ctx.queue, logits, lse, ctx.T, override=ov,
                                          ov_start=start) * dloss
        else:
            dq = ops.infonce_bwd(q_raw, qn, kn, ctx.queue, logits, lse, ctx.T, override=ov, ov_start=start) * dloss
        return dq, None, None, None, None


class InfoNCETensorCoreFn(torch.autograd.Function):
    """Same op as InfoNCEFn on the tcgen05 GEMM: fp16 operands (qn / T, fp16 shadow of the queue), fp32 accumulation.
    The returned logits are a strided view [N, 1+K] of an [N, K+8] buffer (16-byte aligned l_neg block)."""

    @staticmethod
    def forward(ctx, q_raw, k_raw, queue16, T, holder):
        q_raw = q_raw.float().contiguous()
        k_raw = k_raw.float().contiguous()
        qn, kn, buf, lse, loss = ops.infonce_tc_fwd(q_raw, k_raw, queue16, T)
        ctx.save_for_backward(q_raw, qn, kn, buf, lse)
        ctx.queue16, ctx.T, ctx.holder = queue16, T, holder
        ctx.mark_non_differentiable(kn)
        return buf[:, 7:], kn, loss

    @staticmethod
    def backward(ctx, dlogits, _dkn, dloss):
        q_raw, qn, kn, buf, lse = ctx.saved_tensors
        h = ctx.holder or {}
        ov, start = h.get("old"), h.get("start", 0)
        dq = None
        if dlogits is not None:
            dbuf = torch.empty_like(buf)
            dbuf[:, 7:].copy_(dlogits)
            dq = ops.infonce_tc_bwd(q_raw, qn, kn, ctx.queue16, buf, lse, ctx.T, dlogits_buf=dbuf, override=ov,
                                    ov_start=start)
        if dloss is not None:
            d2 = ops.infonce_tc_bwd(q_raw, qn, kn, ctx.queue16, buf, lse, ctx.T, override=ov, ov_start=start) * dloss
            dq = d2 if dq is None else dq + d2
        return dq, None, None, None, None
