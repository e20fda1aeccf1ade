"""Sync-free MF-ViT CA training step (MAIN_CA:859-882 without the four host syncs per iteration).

    trainer = MFViTCATrainer(fusion, vit_cxr, vit_enh, lr=..., momentum=..., weight_decay=...)
    loss = trainer.step(img_cxr, img_enh, target)        # device tensors; returns a 1-element device tensor

One step = grouped encoder forward (mfv_vit_forward, G=2) -> fused CLS cross-attention + heads (mfv_fusion_fwd) ->
cross-entropy on fused + x_cxr + x_enh (mfv_ce_small) -> fusion backward -> encoder backward (mfv_vit_backward) ->
[data-parallel: NCCL all-reduce of the flat gradient buffers] -> fused SGD-momentum step that also rewrites the 16-bit
GEMM shadows (mfv_sgd_step).  No autograd graph, no per-parameter Python loop; parameters stay ordinary nn.Parameters
(views into flat buffers), so state_dict()/checkpoints are unchanged.
"""
import os

import torch
import torch.nn as nn

from . import ops
from ._lib import FusionGrads, MfvError
from .engine import engine_for
from .functions import FUSION_FIELDS


class FlatParams:
    """Packs a list of Parameters into one flat fp32 buffer (Parameters become views) with a matching grad buffer."""

    def __init__(self, params, device):
        self.params = list(params)
        sizes = [(p.numel() + 3) // 4 * 4 for p in self.params]
        self.n = max(sum(sizes), 4)
        self.master = torch.zeros(self.n, device=device, dtype=torch.float32)
        self.grad = torch.zeros(self.n, device=device, dtype=torch.float32)
        self.views, self.gviews = [], []
        off = 0
        for p, sz in zip(self.params, sizes):
            v = self.master[off:off + p.numel()].view(p.shape)
            v.copy_(p.detach().to(device=device, dtype=torch.float32))
            p.data = v
            self.views.append(v)
            self.gviews.append(self.grad[off:off + p.numel()].view(p.shape))
            off += sz

    def adopted(self):
        return all(p.data_ptr() == v.data_ptr() for p, v in zip(self.params, self.views))


class MFViTCATrainer:
    prezero = False     # set from MFVIT_PREZERO in __init__ (class default for objects built without it)
    local_only = False  # True: never all-reduce (bench.py's same-work single-GPU reference inside a data-parallel run)

    def __init__(self, fusion, vit_cxr, vit_enh, lr=1e-3, momentum=0.9, weight_decay=0.0, process_group=None,
                 train_backbones=False, metrics=None, optimizer="sgd", betas=(0.9, 0.999), eps=1e-8):
        """train_backbones=False (default) is the optimisation set of the reference AS WRITTEN: MAIN_CA:435-449 builds
        the optimizer over Fus_CrossViT.parameters() only, i.e. the fusion's own 22 tensors - the two backbones and their
        3-class heads receive gradients (when they require them) but are never stepped; requires_grad
        of every fusion tensor is honoured.  train_backbones=True additionally steps both ViT-S/16 encoders and their
        heads (full fine-tuning: the heavier step bench.py measures, and the one a 173 MB gradient all-reduce implies).
        optimizer: "sgd" (torch.optim.SGD with momentum, MAIN_CA:445-449), "adam" (torch.optim.Adam, L2 weight decay,
        MAIN_CA:453-459) or "adamw" (decoupled decay, MAIN_PRE:339).  The learning rate lives in a 1-element device
        tensor (set_lr), Adam's step count in another, so a captured step follows adjust_learning_rate
        (MAIN_CA:1043-1055) and the bias correction without being captured again."""
        if optimizer not in ("sgd", "adam", "adamw"):
            raise MfvError("optimizer must be 'sgd', 'adam' or 'adamw', not %r" % (optimizer,))
        self.optimizer, self.betas, self.eps = optimizer, tuple(betas), eps
        self._lr_dev = self._step_dev = None
        self._adam_engine = self._adam_small = None
        self.fusion, self.vits = fusion, (vit_cxr, vit_enh)
        self.metrics = metrics      # mfvit.data.EpochMetrics: loss / hits / scores accumulated on the device per step
        self.lr, self.momentum, self.wd = lr, momentum, weight_decay
        self.pg = process_group
        self.train_backbones = train_backbones
        self.engine = engine_for(vit_cxr, vit_enh)
        self.heads = fusion.heads
        self.NC = fusion.num_classes
        self._small = None
        self._bufs = {}
        self.steps = 0
        self._mom_engine = None
        self._mom_small = None
        self.overlap_allreduce = os.environ.get("MFVIT_OVERLAP_ALLREDUCE", "1") != "0"
        # Data parallel: the encoder gradients (173 MB fp32 for both branches) are all-reduced as bf16 - half the volume
        # on NVLink, and half the time the NCCL kernels compete with the backward for SMs; the averaged result is widened
        # back into the fp32 buffer the optimizer reads.  MFVIT_ALLREDUCE=fp32 keeps the fp32 collective (DDP's default).
        self.allreduce_bf16 = os.environ.get("MFVIT_ALLREDUCE", "bf16") == "bf16"
        self._g16 = None
        self._widen = []
        # MFVIT_OVERLAP_OPT=1 (opt-in): the optimizer step of a slice of the encoder runs on a second stream as soon as
        # that slice's gradients are final (the backward then proceeds in block segments), and the zero-fill of the
        # gradient buffer runs beside the forward.  Measured on B200 at 32 pairs: 4.62 ms per step against 4.60 ms with
        # the step at the end - what the overlap hides (0.17 ms of HBM-bound work) is given back by the three extra
        # joins of the weight-gradient side stream at the segment boundaries - so it stays off by default.
        self.overlap_optimizer = os.environ.get("MFVIT_OVERLAP_OPT", "0") == "1"
        # MFVIT_PREZERO=1 (opt-in): the zero-fill of the gradient buffer the split-K weight gradients add into (173 MB,
        # 29 us of HBM writes) runs on a second stream beside the forward instead of in front of the backward.  Measured
        # at 32 pairs: 4.55 ms per step with it, 4.53 without - the forward's epilogues feel the extra HBM writes more
        # than the backward feels the fill - so it stays off.
        self.prezero = os.environ.get("MFVIT_PREZERO", "0") == "1"
        self._opt_stream = None
        self._engine_stepped = False
        self._pending = []
        self._graph = None          # CUDA graph of one whole step (capture_graph)
        self.graph_launches = 0     # kernels of libmfvit.so inside the captured step
        self.graph_replays = 0

    # -- lazily bind to the device of the first batch
    def _prepare(self, device):
        eng = self.engine
        if not eng.is_adopted() or eng.device != device:
            eng.adopt(device)
        if self._small is None or not self._small.adopted():
            self.fusion.to(device)
            params, vh = self.fusion._fusion_params(*self.vits)
            if any(h is None for h in vh):
                raise MfvError("MFViTCATrainer needs nn.Linear backbone heads with num_classes outputs (MAIN_CA:309)")
            for v in self.vits:
                v.head.to(device)
            self._fparams = params  # 26 tensors, FUSION_FIELDS x direction
            self._small = FlatParams(params, device)
            self._pstruct = ops.fusion_param_struct(
                {n: (params[2 * i].data, params[2 * i + 1].data) for i, n in enumerate(FUSION_FIELDS)})
            g = self._small.gviews
            self._gstruct = ops.fusion_param_struct({n: (g[2 * i], g[2 * i + 1]) for i, n in enumerate(FUSION_FIELDS)},
                                                    cls=FusionGrads)
            self._mom_small = torch.zeros_like(self._small.master)
        if self._mom_engine is None or self._mom_engine.device != device:
            self._mom_engine = torch.zeros_like(eng.master)  # SGD momentum buffer / Adam exp_avg
        if self._lr_dev is None or self._lr_dev.device != device:
            self._lr_dev = torch.full((1,), float(self.lr), device=device, dtype=torch.float32)
            self._step_dev = torch.zeros(1, device=device, dtype=torch.int64)
        if self.optimizer != "sgd" and (self._adam_engine is None or self._adam_engine.device != device):
            self._adam_engine = torch.zeros_like(eng.master)      # exp_avg_sq
            self._adam_small = torch.zeros_like(self._small.master)

    def set_lr(self, lr):
        """adjust_learning_rate (MAIN_CA:1043-1055): takes effect at the next step, captured graph or not."""
        self.lr = float(lr)
        if self._lr_dev is not None:
            self._lr_dev.fill_(self.lr)

    def _trainable_ranges(self):
        eng, lay = self.engine, self.engine.layout
        if getattr(self, "_ranges", None) is None:
            out = []
            for g in range(eng.G):
                run = None
                for (name, shape, off), (_, p) in zip(lay.entries, eng._params[g]):
                    n = p.numel()
                    if p.requires_grad:
                        if run is not None and run[1] >= off - 8:  # adjacent up to alignment padding
                            run[1] = off + n
                        else:
                            if run is not None:
                                out.append((g, run[0], (run[1] + 3) // 4 * 4))
                            run = [off, off + n]
                    else:
                        if run is not None:
                            out.append((g, run[0], (run[1] + 3) // 4 * 4))
                            run = None
                if run is not None:
                    out.append((g, run[0], min((run[1] + 3) // 4 * 4, lay.P)))
            self._ranges = out
            self._shadow_complete = False
        return self._ranges

    def forward_backward(self, img_cxr, img_enh, target, reduce_async=False):
        """Forward + backward; gradients land in engine.grads[engine.grad_idx] and self._small.grad.  reduce_async
        (set by step() in data-parallel runs) starts the gradient all-reduce slice by slice during the backward; a
        direct call never issues a collective."""
        device = img_cxr.device
        self._prepare(device)
        eng = self.engine
        B = img_cxr.shape[0]
        lay = eng.layout
        target = target.long()  # MAIN_CA:859 (a no-op for int64 labels)
        enc_grads = eng.any_requires_grad()  # frozen backbones (MAIN_CA:298-305 without --semi-supervised): no backward
        prezero = enc_grads and (self.prezero or (reduce_async and self.overlap_optimizer))
        if prezero:  # the 173 MB zero-fill of the gradient buffer runs beside the forward instead of in front of the backward
            if self._opt_stream is None:
                self._opt_stream = torch.cuda.Stream(device=device)
            self._opt_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._opt_stream):
                ops.fill_(eng.next_grad_buffer().view(-1), 0.0)
        tok, lease = eng.forward([img_cxr, img_enh], save=enc_grads)
        key = (B, device)
        if key not in self._bufs:
            # the loss sums the three heads (MAIN_CA:861), so all three logit gradients are the same [B, NC] tensor: one
            # buffer, d_fused = its first slice, d_x = the other two, filled by ONE broadcast copy per step
            dl3 = torch.empty(3, B, self.NC, device=device, dtype=torch.float32)
            self._bufs[key] = (torch.empty_like(tok), dl3[1:3], ops.fusion_scratch(tok, B, lay.S, lay.C, self.heads), dl3[0], dl3)
        dtok, d_x, scratch, d_fused, dl3 = self._bufs[key]
        ops.fusion_bwd_join(device)  # the previous step's deferred weight-gradient contraction still reads `scratch`
        fused, x = ops.fusion_fwd(tok, self._pstruct, B, lay.S, lay.C, self.heads, self.NC, saved=scratch)
        loss, dlogits = ops.ce_small(fused, x[0], x[1], target)
        ops.fill_(self._small.grad, 0.0)
        dl3.copy_(dlogits.unsqueeze(0).expand_as(dl3))
        # only dtok is needed by the encoder backward: the fusion's own parameter gradients are contracted on the
        # library's side stream meanwhile and joined below (all buffers they read are owned by the trainer)
        ops.fusion_bwd(tok, self._pstruct, self._gstruct, d_fused, d_x, B, lay.S, lay.C, self.heads, self.NC, dtok=dtok,
                       scratch=scratch, defer=True)
        self._pending = []
        self._engine_stepped = False
        dist = torch.distributed
        overlap_ar = reduce_async and self._overlap_allreduce()
        needs_ar = (not self.local_only and dist.is_available() and dist.is_initialized()
                    and dist.get_world_size(self.pg) > 1)
        # slices may only be stepped early when their gradients are final: one GPU, or all-reduced slice by slice
        overlap_opt = (reduce_async and enc_grads and self.train_backbones and self.overlap_optimizer
                       and (overlap_ar or not needs_ar))
        if not enc_grads:
            ops.fusion_bwd_join(device)
            grad = None
            if overlap_ar:
                self._pending.append(dist.all_reduce(self._small.grad, op=dist.ReduceOp.AVG, group=self.pg, async_op=True))
        elif overlap_ar or overlap_opt:
            # The encoder backward runs in MFVIT_DP_SEGMENTS (default three) block segments.  Blocks are contiguous in
            # the flat layout, so the gradient slice a segment finished is one contiguous range per branch: data parallel,
            # it is all-reduced (NCCL, asynchronously) while the next segment computes; then - or at once on one GPU -
            # the optimizer steps that slice on its own stream.
            d = lay.depth
            # Cut points (block indices, descending): three equal segments by default; MFVIT_DP_SEGMENTS=n asks for n,
            # MFVIT_DP_CUTS="6,2" sets them explicitly.  Shrinking the last segment (the one whose all-reduce nothing
            # hides) was measured at 2 GPUs and is within the run-to-run spread: 8.49 (thirds) / 8.53 (6,2) / 8.45 (6,1) /
            # 8.54 (5,1) / 8.50 ms (7,3,1) - what stays exposed (~0.3 ms) is not that segment's bytes.
            env_cuts, env_nseg = os.environ.get("MFVIT_DP_CUTS", ""), os.environ.get("MFVIT_DP_SEGMENTS", "")
            if env_cuts:
                inner = {min(max(int(c), 0), d) for c in env_cuts.split(",") if c.strip()}
            else:
                nseg = max(1, int(env_nseg)) if env_nseg else 3
                inner = {(d * k) // nseg for k in range(nseg + 1)}
            cuts = sorted(inner | {0, d}, reverse=True)
            segments = [(cuts[i] - 1, cuts[i + 1]) for i in range(len(cuts) - 1)]
            main = torch.cuda.current_stream()
            if prezero:
                main.wait_stream(self._opt_stream)
            if overlap_opt and self.optimizer != "sgd":
                self._step_dev.add_(1)

            if overlap_ar and self.allreduce_bf16 and (self._g16 is None or self._g16.device != device):
                self._g16 = torch.empty(eng.G, lay.P, device=device, dtype=torch.bfloat16)
            self._widen = []

            def on_slice(grad, lo, hi):
                works = []
                if overlap_ar:
                    assert lo % 8 == 0 and hi % 8 == 0  # block boundaries of the flat layout: 16-byte aligned 16-bit slices
                    lo8 = lo
                    for g in range(eng.G):
                        if self.allreduce_bf16:
                            ops.cast_shadow(grad[g, lo8:hi], self._g16[g, lo8:hi], None)
                            works.append(dist.all_reduce(self._g16[g, lo8:hi], op=dist.ReduceOp.AVG, group=self.pg,
                                                         async_op=True))
                            self._widen.append((self._g16[g, lo8:hi], grad[g, lo8:hi]))
                        else:
                            works.append(dist.all_reduce(grad[g, lo:hi], op=dist.ReduceOp.AVG, group=self.pg, async_op=True))
                if overlap_opt:
                    self._opt_stream.wait_stream(main)
                    with torch.cuda.stream(self._opt_stream):
                        for w in works:
                            w.wait()
                        for src, dst in self._widen:
                            ops.cast_bf16_f32_(src, dst)
                        self._widen = []
                        self._step_engine(grad, lo, hi)
                else:
                    self._pending.extend(works)
            grad = eng.backward(lease, dtok, zero=not prezero, segments=segments, on_segment=on_slice)
            ops.fusion_bwd_join(device)
            if overlap_opt:
                self._engine_stepped = True
            if overlap_ar:
                self._pending.append(dist.all_reduce(self._small.grad, op=dist.ReduceOp.AVG, group=self.pg, async_op=True))
        else:
            if prezero:
                torch.cuda.current_stream().wait_stream(self._opt_stream)
            grad = eng.backward(lease, dtok, zero=not prezero)
            ops.fusion_bwd_join(device)
        self._last = (fused, x)
        return loss, grad

    def _overlap_allreduce(self):
        dist = torch.distributed
        return (not self.local_only and dist.is_available() and dist.is_initialized() and dist.get_world_size(self.pg) > 1
                and dist.get_backend(self.pg) == "nccl" and self.overlap_allreduce)

    def all_reduce(self, grad):
        if getattr(self, "_pending", None):
            for w in self._pending:  # stream-level waits: the optimizer step is ordered after the NCCL kernels
                w.wait()
            self._pending = []
            for src, dst in self._widen:  # bf16 all-reduce: widen the averaged slices back into the fp32 gradient buffer
                ops.cast_bf16_f32_(src, dst)
            self._widen = []
            return
        if self.local_only:
            return
        if self.pg is None and not (torch.distributed.is_available() and torch.distributed.is_initialized()):
            return
        ws = torch.distributed.get_world_size(self.pg)
        if ws == 1:
            return
        # mean over ranks (DDP semantics).  NCCL averages inside the collective (no extra pass over the 173 MB flat
        # gradient buffer); gloo (CPU tests) has no AVG, so pre-divide there.
        avg = torch.distributed.get_backend(self.pg) == "nccl"
        for t in ((self._small.grad,) if grad is None else (grad, self._small.grad)):
            if avg:
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.AVG, group=self.pg)
            else:
                t.div_(ws)
                torch.distributed.all_reduce(t, group=self.pg)

    def _small_ranges(self):
        """Element ranges of the packed fusion / head buffer the optimizer steps: tensors that require grad, and - in
        the reference's as-written mode - not the backbone heads (vhead_*, the last four tensors)."""
        if getattr(self, "_sranges", None) is None:
            fp = self._small
            n_t = len(fp.params) if self.train_backbones else len(fp.params) - 4
            out, off, run = [], 0, None
            for i, p in enumerate(fp.params):
                sz = (p.numel() + 3) // 4 * 4
                if i < n_t and p.requires_grad:
                    if run is not None and run[1] == off:
                        run[1] = off + sz
                    else:
                        if run is not None:
                            out.append(tuple(run))
                        run = [off, off + sz]
                off += sz
            if run is not None:
                out.append(tuple(run))
            self._sranges = out
        return self._sranges

    def _update(self, p, g, m, v, shadow, shadow16):
        if self.optimizer != "sgd":
            ops.adam_step_dev_(p, g, m, v, shadow, self._lr_dev, self.betas, self.eps, self.wd,
                               self.optimizer == "adamw", self._step_dev, shadow16=shadow16)
        else:
            ops.sgd_step_dev_(p, g, m, shadow, self._lr_dev, self.momentum, self.wd, self.steps == 0, shadow16=shadow16)

    def _step_engine(self, grad, lo_elem=0, hi_elem=None):
        """Optimizer step of the encoder parameters inside [lo_elem, hi_elem) of every branch's flat buffer."""
        eng = self.engine
        adam = self.optimizer != "sgd"
        hi_elem = eng.layout.P if hi_elem is None else hi_elem
        # contiguous runs of trainable tensors (pos_embed is a fixed table; stop_grad_conv1 freezes the conv): frozen
        # ranges must not see weight decay, so they are skipped rather than stepped with a zero gradient
        for g, lo, hi in self._trainable_ranges():
            lo, hi = max(lo, lo_elem), min(hi, hi_elem)
            if lo >= hi:
                continue
            sl = slice(lo, hi)
            self._update(eng.master[g, sl], grad[g, sl], self._mom_engine[g, sl],
                         self._adam_engine[g, sl] if adam else None, eng.shadow[g, sl],
                         eng.shadow16[g, sl] if eng.fwd_f16 else None)

    def optimizer_step(self, grad):
        eng = self.engine
        adam = self.optimizer != "sgd"
        if adam and not self._engine_stepped:
            self._step_dev.add_(1)  # 1-based update count, read by the kernels on the device
        if self.train_backbones:
            if self._engine_stepped:  # slices were stepped on the optimizer stream during the backward: join it
                torch.cuda.current_stream().wait_stream(self._opt_stream)
                self._engine_stepped = False
            else:
                self._step_engine(grad)
            if not self._shadow_complete:
                eng.cast_shadow()  # frozen ranges, once
                self._shadow_complete = True
            eng.mark_shadow_fresh()  # the step rewrote the GEMM shadows: the next forward skips the cast pass
        else:
            eng.mark_shadow_fresh()  # nobody steps the encoders here: the shadows the forward cast are still right
        adam_small = self._adam_small if adam else None
        for lo, hi in self._small_ranges():
            sl = slice(lo, hi)
            self._update(self._small.master[sl], self._small.grad[sl], self._mom_small[sl],
                         adam_small[sl] if adam else None, None, None)
        self.steps += 1

    def _step_eager(self, img_cxr, img_enh, target):
        loss, grad = self.forward_backward(img_cxr, img_enh, target, reduce_async=True)
        if self.metrics is not None:  # MAIN_CA:884-899 without the host round trips
            self.metrics.accumulate(*self.logits(), target, loss)
        self.all_reduce(grad)
        self.optimizer_step(grad)
        return loss

    def step(self, img_cxr, img_enh, target):
        if self._graph is not None and tuple(img_cxr.shape) == tuple(self._g_inputs[0].shape):
            for dst, src in zip(self._g_inputs, (img_cxr, img_enh, target)):
                if dst.data_ptr() != src.data_ptr():
                    dst.copy_(src, non_blocking=True)
            eng = self.engine
            if eng._shadow_sig != eng._param_sig():
                # the captured step holds no cast pass (its own optimizer rewrites the shadows): parameters edited from
                # outside since the last replay (load_state_dict, p.copy_) are cast here, once
                eng.cast_shadow()
                eng.mark_shadow_fresh()
            self._graph.replay()
            self.steps += 1
            self.graph_replays += 1
            return self._g_loss
        return self._step_eager(img_cxr, img_enh, target)

    def capture_graph(self, img_cxr, img_enh, target):
        """Capture one whole step (forward, loss, backward incl. the side-stream weight gradients, optimizer) into a
        CUDA graph replayed by step(): ~240 kernel launches become one graph launch, so a per-step host read of the loss
        (MAIN_CA:884) no longer starves the GPU.  Parameters and optimizer state are left exactly as they were: the
        two eager warm-up steps CUDA graph capture needs are undone from a snapshot."""
        # Data parallel: the NCCL all-reduce is captured with the step (every rank must capture and replay in lock step).
        from . import _lib
        device = img_cxr.device
        self._prepare(device)
        lib = _lib.init(device.index if device.index is not None else torch.cuda.current_device())
        eng = self.engine
        state = [eng.master, self._mom_engine, self._small.master, self._mom_small, self._step_dev]
        if self.optimizer != "sgd":
            state += [self._adam_engine, self._adam_small]
        snap = [t.clone() for t in state]
        steps0 = self.steps
        self._g_inputs = [torch.empty_like(t) for t in (img_cxr, img_enh, target)]
        for dst, src in zip(self._g_inputs, (img_cxr, img_enh, target)):
            dst.copy_(src)
        # (Capturing on a high-priority stream - so that pending CTAs of the critical chain are placed before those of the
        # weight-gradient side stream - was measured and is worse: 4.85 against 4.58 ms per step; the starved side stream
        # piles its GEMMs up behind the backward.  MFVIT_MAIN_PRIORITY=1 reproduces it.)
        prio = -1 if os.environ.get("MFVIT_MAIN_PRIORITY", "0") == "1" else 0
        side = torch.cuda.Stream(device=device, priority=prio)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):  # allocations, cudaFuncSetAttribute, side-stream / event creation happen here
                self._step_eager(*self._g_inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        n0 = lib.mfv_launch_count()
        with torch.cuda.graph(graph, stream=side):
            self._g_loss = self._step_eager(*self._g_inputs)
        self.graph_launches = int(lib.mfv_launch_count() - n0)
        # undo the warm-up steps (capture itself executes nothing); the 16-bit GEMM shadows are rebuilt from the master
        for dst, src in zip(state, snap):
            dst.copy_(src)
        eng.cast_shadow()
        if self.metrics is not None:
            self.metrics.reset()  # the warm-up steps are not part of the epoch
        self.steps = max(steps0, 1)  # the captured step is a steady-state one (momentum buffers exist)
        eng.mark_shadow_fresh()
        self._graph = graph
        return self

    def logits(self):
        """(fused, x_cxr, x_enh) of the most recent step."""
        fused, x = self._last
        return fused, x[0], x[1]
