"""Restatement of the MoCo-v3-structure / v2-loss step logic (BLD = moco/builder_vit_mocov3structure_mocov2loss.py),
pure PyTorch, single-process or torch.distributed (gloo/nccl).  Functions cite the BLD lines they follow."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def ema_update(params_k, params_q, m):
    """BLD:83-89.  `m` is a Python float; torch rounds m and (1.-m) to fp32 separately."""
    for pq, pk in zip(params_q, params_k):
        pk.data = pk.data * m + pq.data * (1.0 - m)


def build_mlp(num_layers, input_dim, mlp_dim, output_dim, last_bn=True):
    """BLD:62-78."""
    mlp = []
    for l in range(num_layers):
        dim1 = input_dim if l == 0 else mlp_dim
        dim2 = output_dim if l == num_layers - 1 else mlp_dim
        mlp.append(nn.Linear(dim1, dim2, bias=False))
        if l < num_layers - 1:
            mlp.append(nn.BatchNorm1d(dim2))
            mlp.append(nn.ReLU(inplace=True))
        elif last_bn:
            mlp.append(nn.BatchNorm1d(dim2, affine=False))
    return nn.Sequential(*mlp)


def infonce_logits(q_raw, k_raw, queue, T):
    """BLD:165,175,183-194: normalise, l_pos, l_neg, cat, /T, labels 0."""
    q = F.normalize(q_raw, dim=1)
    k = F.normalize(k_raw, dim=1)
    l_pos = torch.einsum("nc,nc->n", [q, k]).unsqueeze(-1)
    l_neg = torch.einsum("nc,ck->nk", [q, queue.clone().detach()])
    logits = torch.cat([l_pos, l_neg], dim=1)
    logits /= T
    labels = torch.zeros(logits.shape[0], dtype=torch.long, device=logits.device)
    return logits, labels, q, k


def concat_all_gather(tensor):
    """BLD:229-240 (rank-major concat)."""
    if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
        return tensor
    ws = torch.distributed.get_world_size()
    out = [torch.ones_like(tensor) for _ in range(ws)]
    torch.distributed.all_gather(out, tensor, async_op=False)
    return torch.cat(out, dim=0)


def dequeue_and_enqueue(queue, queue_ptr, keys, K):
    """BLD:91-105."""
    keys = concat_all_gather(keys)
    bs = keys.shape[0]
    ptr = int(queue_ptr)
    assert K % bs == 0
    queue[:, ptr:ptr + bs] = keys.t()
    queue_ptr[0] = (ptr + bs) % K


def cosine_momentum(epoch_float, total_epochs, m0):
    """MAIN_PRE:626-629."""
    import math
    return 1.0 - 0.5 * (1.0 + math.cos(math.pi * epoch_float / total_epochs)) * (1.0 - m0)


class MoCoViT(nn.Module):
    """The whole model of BLD:16-60,215-225 and its forward BLD:154-199, assembled from the functions above over the
    oracle ViT (oracle/vit_ref.py).  Test infrastructure: the module tree and state-dict keys equal the reference's, so
    weights move between this class, the real reference class (tests/golden/moco_ref.pt pins the pieces against it) and
    the drop-in with load_state_dict(strict=True)."""

    def __init__(self, base_encoder, dim=256, mlp_dim=4096, T=1.0, K=65536):
        super().__init__()
        self.T, self.K = T, K
        self.base_encoder = base_encoder(num_classes=mlp_dim)                    # BLD:28-30
        self.momentum_encoder = base_encoder(num_classes=mlp_dim)
        hidden = self.base_encoder.head.weight.shape[1]                          # BLD:216-225
        del self.base_encoder.head, self.momentum_encoder.head
        self.base_encoder.head = build_mlp(3, hidden, mlp_dim, dim)
        self.momentum_encoder.head = build_mlp(3, hidden, mlp_dim, dim)
        self.predictor = build_mlp(2, dim, mlp_dim, dim)
        for pb, pm in zip(self.base_encoder.parameters(), self.momentum_encoder.parameters()):   # BLD:48-52
            pm.data.copy_(pb.data)
            pm.requires_grad = False
        self.register_buffer("queue", F.normalize(torch.randn(dim, K), dim=0))  # BLD:55-57
        self.register_buffer("queue_ptr", torch.zeros(1, dtype=torch.long))

    def forward(self, im_q, im_k, m):
        q = self.predictor(self.base_encoder(im_q))                              # BLD:163-164
        with torch.no_grad():                                                    # BLD:168-181 (shuffle: identity at 1 rank)
            ema_update(list(self.momentum_encoder.parameters()), list(self.base_encoder.parameters()), m)
            k = self.predictor(self.momentum_encoder(im_k))
        logits, labels, _, kn = infonce_logits(q, k, self.queue, self.T)          # BLD:165,175,183-194
        dequeue_and_enqueue(self.queue, self.queue_ptr, kn.detach(), self.K)     # BLD:197
        return logits, labels
