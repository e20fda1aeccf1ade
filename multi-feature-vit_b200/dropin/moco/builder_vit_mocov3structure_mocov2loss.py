"""Drop-in for moco/builder_vit_mocov3structure_mocov2loss.py (BLD): MoCo with the v3 structure (3-layer projector in
`head`, 2-layer predictor, momentum encoder) and the v2 loss (65 536-entry queue, l_pos/l_neg logits, label 0).

Same class / attribute / buffer names and forward signature as the reference (BLD:11-226); the hot ops are kernels:
  * both ViT-S/16 encoders            -> mfvit.engine (mfv_vit_forward / mfv_vit_backward)
  * _momentum_update_key_encoder      -> mfv_ema_update, one multi-tensor launch, bit-exact with BLD:88
  * normalize + l_pos + l_neg + cat/T -> mfv_infonce_tc_fwd / _bwd: l_neg on the tcgen05 GEMM against an fp16 shadow of
                                         the queue (no queue.clone(), BLD:185); MFVIT_INFONCE=fp32 selects the exact
                                         fp32 FMA kernels mfv_infonce_fwd / _bwd
  * _dequeue_and_enqueue              -> one NCCL all_gather_into_tensor + mfv_enqueue_keys (transposed write)
The projector / predictor MLPs (Linear-BN-ReLU, 0.4 % of the step FLOPs) remain torch.nn modules so that
SyncBatchNorm conversion and DDP wrapping (MAIN_PRE:297,312) keep working unchanged.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _path  # noqa: E402,F401
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
from mfvit import ops  # noqa: E402
from mfvit.functions import InfoNCEFn, InfoNCETensorCoreFn  # noqa: E402


def _dist_on():
    return torch.distributed.is_available() and torch.distributed.is_initialized()


class MoCo(nn.Module):
    """Build a MoCo model with a base encoder, a momentum encoder, and two MLPs (BLD:11-60)."""
    predictor_on_keys = True  # False in builder_vit_mocov3structure_mocov2loss_noprediction_q

    def __init__(self, base_encoder, args, dim=256, mlp_dim=4096, T=1.0):
        super(MoCo, self).__init__()
        self.T = T
        self.K = 65536
        if not args.arch.startswith('vit'):
            raise NotImplementedError("only the ViT path of MoCo is on the MF-ViT hot path (SURVEY 2, rows 7/13)")
        self.base_encoder = base_encoder(num_classes=mlp_dim)
        self.momentum_encoder = base_encoder(num_classes=mlp_dim)
        self._build_projector_and_predictor_mlps(dim, mlp_dim)
        # BLD:48-52: the momentum encoder starts as a copy of the base encoder and never sees a gradient
        with torch.no_grad():
            for src, dst in zip(self.base_encoder.parameters(), self.momentum_encoder.parameters()):
                dst.copy_(src)
                dst.requires_grad_(False)
        self.register_buffer("queue", torch.randn(dim, self.K))
        self.queue = nn.functional.normalize(self.queue, dim=0)
        self.register_buffer("queue_ptr", torch.zeros(1, dtype=torch.long))
        self._ema_cache = None
        self._ptr_host = None
        # fp16 shadow of the queue for the tensor-core InfoNCE (not a buffer: it never reaches a state dict)
        self._queue16 = None
        self._queue16_sig = None
        self.infonce_precision = os.environ.get("MFVIT_INFONCE", "tc")
        # True: reproduce BLD:107-152 (all-gather the key images, global shuffle) even when it cannot change the result
        self.force_batch_shuffle = False

    def _build_mlp(self, num_layers, input_dim, mlp_dim, output_dim, last_bn=True):
        """BLD:62-78: bias-free Linear layers; hidden ones followed by BN + ReLU, the last one by a BN without affine
        parameters when `last_bn`.  Module indices (= state-dict keys) match the reference's nn.Sequential."""
        widths = [input_dim] + [mlp_dim] * (num_layers - 1) + [output_dim]
        layers = []
        for i, (fan_in, fan_out) in enumerate(zip(widths[:-1], widths[1:])):
            layers.append(nn.Linear(fan_in, fan_out, bias=False))
            if i + 1 < num_layers:
                layers += [nn.BatchNorm1d(fan_out), nn.ReLU(inplace=True)]
            elif last_bn:
                layers.append(nn.BatchNorm1d(fan_out, affine=False))
        return nn.Sequential(*layers)

    def _build_projector_and_predictor_mlps(self, dim, mlp_dim):
        pass

    # ------------------------------------------------------------------------------------------------ EMA (BLD:83-89)
    @torch.no_grad()
    def _momentum_update_key_encoder(self, m):
        pq = list(self.base_encoder.parameters())
        pk = list(self.momentum_encoder.parameters())
        if not pq[0].is_cuda:
            raise ops.MfvError("MoCo momentum update runs only on CUDA sm_100a devices (no CPU fallback)")
        # encoder parameters live in the engines' flat buffers: adopt first so the chunk table sees final pointers
        for enc in (self.base_encoder, self.momentum_encoder):
            from mfvit.engine import engine_for
            eng = engine_for(enc)
            if not eng.is_adopted() or eng.device != pq[0].device:
                eng.adopt(pq[0].device)
        sig = tuple(p.data_ptr() for p in pq) + tuple(p.data_ptr() for p in pk)
        if self._ema_cache is None or self._ema_cache[0] != sig:
            chunks, n, mx = ops.make_ema_chunks([(k.data, q.data) for q, k in zip(pq, pk)], pq[0].device)
            self._ema_cache = (sig, chunks, n, mx)
        _, chunks, n, mx = self._ema_cache
        ops.ema_update_(chunks, n, mx, m)
        from mfvit.engine import engine_for
        engine_for(self.momentum_encoder).invalidate_shadow()

    # ------------------------------------------------------------------------------------------------ queue (BLD:91-105)
    @torch.no_grad()
    def _dequeue_and_enqueue(self, keys, holder=None):
        keys = concat_all_gather(keys)
        batch_size = keys.shape[0]
        if self._ptr_host is None:
            self._ptr_host = int(self.queue_ptr)  # one sync at the first step only (the reference syncs every step)
        ptr = self._ptr_host
        assert self.K % batch_size == 0  # BLD:99: the queue length must be a multiple of the global batch
        if holder is not None:
            holder["start"] = ptr
            holder["old"] = self.queue[:, ptr:ptr + batch_size].clone()
        ops.enqueue_keys_(keys.float().contiguous(), self.queue, ptr)
        if self._queue16 is not None and self._queue16_sig == self._queue_sig():
            ops.queue16_update_(self.queue, self._queue16, ptr, batch_size)  # keep the fp16 shadow current
        ptr = (ptr + batch_size) % self.K
        self._ptr_host = ptr
        self.queue_ptr.fill_(ptr)

    def _load_from_state_dict(self, *a, **k):
        self._ptr_host = None
        self._queue16_sig = None
        return super()._load_from_state_dict(*a, **k)

    def _queue_sig(self):
        # in-place edits through torch bump _version; .to()/.cuda() change the storage; our own enqueue does neither
        return (self.queue.data_ptr(), self.queue._version, str(self.queue.device))

    def _queue16_current(self):
        """fp16 shadow of `queue`, rebuilt in one pass whenever the fp32 queue was replaced or edited behind our back."""
        sig = self._queue_sig()
        if self._queue16 is None or self._queue16_sig != sig or self._queue16.device != self.queue.device:
            self._queue16 = torch.empty(self.queue.shape, device=self.queue.device, dtype=torch.float16)
            ops.queue16_update_(self.queue, self._queue16)
            self._queue16_sig = sig
        return self._queue16

    # ------------------------------------------------------------------------------------------------ shuffle (BLD:107-152)
    @staticmethod
    def _rank_world():
        if _dist_on():
            return torch.distributed.get_rank(), torch.distributed.get_world_size()
        return 0, 1

    @torch.no_grad()
    def _batch_shuffle_ddp(self, x):
        """BLD:107-135: every rank takes its slice of one global permutation (drawn on each rank, rank 0's wins) of
        the all-gathered batch; also returns the inverse permutation for _batch_unshuffle_ddp."""
        rank, world = self._rank_world()
        everything = concat_all_gather(x)
        perm = torch.randperm(everything.shape[0]).to(x.device)  # CPU generator, as the reference draws it
        if world > 1:
            torch.distributed.broadcast(perm, src=0)
        inverse = torch.argsort(perm)
        mine = perm.chunk(world)[rank]
        return everything[mine], inverse

    @torch.no_grad()
    def _batch_unshuffle_ddp(self, x, idx_unshuffle):
        """BLD:137-152: gather the shuffled features and pick this rank's rows in the original order."""
        rank, world = self._rank_world()
        return concat_all_gather(x)[idx_unshuffle.chunk(world)[rank]]

    def _shuffle_is_noop(self):
        """Shuffle-BN only changes which samples share BatchNorm statistics.  With one rank, or with SyncBatchNorm
        (MAIN_PRE:297: statistics are global), the key features are the same function of the key images."""
        if self.force_batch_shuffle:
            return False
        if not _dist_on() or torch.distributed.get_world_size() == 1:
            return True
        bns = [m for m in list(self.momentum_encoder.modules()) + list(self.predictor.modules())
               if isinstance(m, nn.modules.batchnorm._BatchNorm)]
        return all(isinstance(m, nn.SyncBatchNorm) for m in bns)

    # ------------------------------------------------------------------------------------------------ forward (BLD:154-199)
    def forward(self, im_q, im_k, m):
        q = self.predictor(self.base_encoder(im_q))  # queries: NxC (normalised inside the InfoNCE kernel)
        with torch.no_grad():
            self._momentum_update_key_encoder(m)
            # BLD:174 runs the keys through the predictor too; the `_noprediction_q` builder (BLD_NOPRED:175) does not
            key_head = self.predictor if self.predictor_on_keys else (lambda t: t)
            if self._shuffle_is_noop():
                k = key_head(self.momentum_encoder(im_k))
            else:
                im_k, idx_unshuffle = self._batch_shuffle_ddp(im_k)
                k = key_head(self.momentum_encoder(im_k))
                k = self._batch_unshuffle_ddp(k, idx_unshuffle)
        holder = {}
        if self.infonce_precision == "fp32":
            logits, kn, _ = InfoNCEFn.apply(q, k, self.queue, self.T, holder)
        else:
            logits, kn, _ = InfoNCETensorCoreFn.apply(q, k, self._queue16_current(), self.T, holder)
        labels = torch.zeros(logits.shape[0], dtype=torch.long, device=logits.device)
        self._dequeue_and_enqueue(kn, holder)
        return logits, labels


class MoCo_ResNet(MoCo):
    def __init__(self, *a, **k):
        raise NotImplementedError("MoCo_ResNet (CNN path) is outside the MF-ViT hot path (SURVEY 2, row 13)")


class MoCo_ViT(MoCo):
    def _build_projector_and_predictor_mlps(self, dim, mlp_dim):  # BLD:216-225
        hidden_dim = self.base_encoder.head.weight.shape[1]
        del self.base_encoder.head, self.momentum_encoder.head  # remove original fc layer
        self.base_encoder.head = self._build_mlp(3, hidden_dim, mlp_dim, dim)
        self.momentum_encoder.head = self._build_mlp(3, hidden_dim, mlp_dim, dim)
        self.predictor = self._build_mlp(2, dim, mlp_dim, dim)


@torch.no_grad()
def concat_all_gather(tensor):
    """Rank-major concatenation of `tensor` from every rank (BLD:229-240): one all_gather_into_tensor into a single
    preallocated buffer instead of world_size ones_like buffers + torch.cat.  No gradient."""
    if not _dist_on() or torch.distributed.get_world_size() == 1:
        return tensor
    ws = torch.distributed.get_world_size()
    tensor = tensor.contiguous()
    out = torch.empty((ws * tensor.shape[0],) + tuple(tensor.shape[1:]), dtype=tensor.dtype, device=tensor.device)
    torch.distributed.all_gather_into_tensor(out, tensor)
    return out
