"""ViT-S/16 encoder engine: flat parameter storage, activation workspaces and the two calls into the native runtime
(mfv_vit_forward / mfv_vit_backward).  One engine drives G branches (G=2: the CXR and the enhanced-CXR backbones of
MF-ViT CA run as grouped launches with per-group weights; G=1: a single ViT, e.g. the MoCo base / momentum encoder).

HBM layout (per group g, P = padded parameter count):
  master f32 [G][P]  - the nn.Parameters of the drop-in modules are *views* into this buffer (state-dict keys unchanged)
  shadow bf16 [G][P] - GEMM operand copy, refreshed by one cast kernel (or by the fused optimiser step)
  grad   f32 [G][P]  - weight gradients (split-K red.add target); two buffers ping-pong so autograd may keep a view
"""
import ctypes as C
import operator
import os
import weakref

import torch

from . import _lib, ops
from ._lib import MfvError, VitPlan, check

_BLOCK_FIELDS = [
    ("norm1.weight", "r_ln1_w"), ("norm1.bias", "r_ln1_b"), ("attn.qkv.weight", "r_qkv_w"),
    ("attn.qkv.bias", "r_qkv_b"), ("attn.proj.weight", "r_proj_w"), ("attn.proj.bias", "r_proj_b"),
    ("norm2.weight", "r_ln2_w"), ("norm2.bias", "r_ln2_b"), ("mlp.fc1.weight", "r_fc1_w"), ("mlp.fc1.bias", "r_fc1_b"),
    ("mlp.fc2.weight", "r_fc2_w"), ("mlp.fc2.bias", "r_fc2_b"),
]


# Forward GEMM operand format.  "fp16": IEEE half operands in the forward (8x finer than bf16 -> logits within 2e-3 of the
# fp32 reference), bf16 in the backward (range-safe gradients).  "bf16": bf16 everywhere (no dual-format activations).
FWD_PRECISION = os.environ.get("MFVIT_FWD_PRECISION", "fp16")
FP16_SAFE_BOUND = 3.0e4     # half of the largest finite fp16 value: headroom for the GEMMs' own rounding
FP16_CHECK_EVERY = 64       # shadow casts between two range checks (weights move slowly; the bound has 10-100x slack)


_VERSION_OF = operator.attrgetter("_version")


def _align8(n):
    return (n + 7) // 8 * 8


class ViTLayout:
    """Offsets (elements) of every encoder tensor inside one group's flat buffer; names are the timm state-dict keys."""

    def __init__(self, img_size, patch, C, depth, heads, hidden):
        if patch != 16:
            raise MfvError("mfvit kernels are specialised for 16x16 patches")
        self.img, self.C, self.depth, self.H, self.hidden = img_size, C, depth, heads, hidden
        self.np = (img_size // patch) ** 2
        self.S = self.np + 1
        self.entries = []  # (name, shape, offset)
        off = 0

        def add(name, shape):
            nonlocal off
            n = 1
            for s in shape:
                n *= s
            self.entries.append((name, tuple(shape), off))
            off = _align8(off + n)

        add("cls_token", (1, 1, C))
        add("pos_embed", (1, self.S, C))
        add("patch_embed.proj.weight", (C, 3, patch, patch))
        add("patch_embed.proj.bias", (C,))
        shapes = {"norm1.weight": (C,), "norm1.bias": (C,), "attn.qkv.weight": (3 * C, C), "attn.qkv.bias": (3 * C,),
                  "attn.proj.weight": (C, C), "attn.proj.bias": (C,), "norm2.weight": (C,), "norm2.bias": (C,),
                  "mlp.fc1.weight": (hidden, C), "mlp.fc1.bias": (hidden,), "mlp.fc2.weight": (C, hidden),
                  "mlp.fc2.bias": (C,)}
        self.off_block0 = off
        for i in range(depth):
            start = off
            for nm, _ in _BLOCK_FIELDS:
                add("blocks.%d.%s" % (i, nm), shapes[nm])
            if i == 0:
                self.block_stride = off - start
        add("norm.weight", (C,))
        add("norm.bias", (C,))
        self.P = _align8(off)
        self.offset = {n: o for n, _, o in self.entries}
        self.shape = {n: s for n, s, _ in self.entries}

    def key(self):
        return (self.img, self.C, self.depth, self.H, self.hidden)


class _Workspace:
    """All activation / gradient scratch for one (B, save) configuration, carved from a few big allocations."""

    def __init__(self, lay, G, B, save, device, fwd_f16):
        M = B * lay.S
        C_, Hd = lay.C, lay.hidden
        nblk = lay.depth if save else 1
        nx = 2 * lay.depth + 1 if save else 2
        nxn = 2 * lay.depth if save else 1
        nst = 2 * lay.depth + 1 if save else 1
        bf, f32 = torch.bfloat16, torch.float32

        def e(n, dt):
            return torch.empty(int(n), device=device, dtype=dt)

        self.G, self.B, self.save, self.busy = G, B, save, False
        self.patches = e(G * B * lay.np * 768, bf)
        self.acc = e(G * B * lay.np * C_, f32)
        self.x = e(nx * G * M * C_, f32)
        self.xn = e(nxn * G * M * C_, bf)
        self.stats = e(nst * 2 * G * M, f32)
        self.qkv = e(nblk * G * M * 3 * C_, bf)
        self.attn_o = e(nblk * G * M * C_, bf)
        self.lse = e(nblk * G * B * lay.H * lay.S, f32)
        self.u = e(nblk * G * M * Hd, bf)
        self.gact = e(nblk * G * M * Hd, bf)
        # bf16 copies of the fp16 forward operands, read by the backward GEMMs (fp16-forward mode, training only)
        dual = fwd_f16 and save
        self.patches_bf = e(G * B * lay.np * 768, bf) if dual else None
        self.xn_bf = e(nxn * G * M * C_, bf) if dual else None
        self.attn_o_bf = e(nblk * G * M * C_, bf) if dual else None
        # bf16 gelu(u) for the fc2 weight gradient: one slot per block, stored by the fc1 epilogue (default), or - with
        # MFVIT_GELU_TWIN=0 - one slot that the fc2-dgrad epilogue recomputes per block (27 MB per pair less memory, 0.3 % slower)
        self.gact_twin = bool(dual) and os.environ.get("MFVIT_GELU_TWIN", "1") != "0"
        self.gact_bf = e((nblk if self.gact_twin else 1) * G * M * Hd, bf) if dual else None
        self.bwd = None

    def ensure_bwd(self, lay, device):
        if self.bwd is None:
            G, B = self.G, self.B
            M = B * lay.S
            bf, f32 = torch.bfloat16, torch.float32
            self.bwd = dict(
                dx0=torch.empty(G * M * lay.C, device=device, dtype=f32),
                dx1=torch.empty(G * M * lay.C, device=device, dtype=f32),
                dx16_0=torch.empty(G * M * lay.C, device=device, dtype=bf),
                dx16_1=torch.empty(G * M * lay.C, device=device, dtype=bf),
                dhid=torch.empty(G * M * lay.hidden, device=device, dtype=bf),
                dxn=torch.empty(G * M * lay.C, device=device, dtype=bf),
                d_o=torch.empty(G * M * lay.C, device=device, dtype=bf),
                dqkv=torch.empty(G * M * 3 * lay.C, device=device, dtype=bf),
                delta=torch.empty(G * B * lay.H * lay.S, device=device, dtype=f32),
                dacc=torch.empty(G * B * lay.np * lay.C, device=device, dtype=bf),
                # fp32 dQ accumulator of the long-sequence (S > 224) tcgen05 attention backward
                attn_ws=(torch.empty(G * M * lay.C, device=device, dtype=f32)
                         if (lay.S > 224 and lay.C // lay.H == 64) else None),
            )
        return self.bwd


class _Lease:
    """Marks a workspace busy between a saving forward and the end of its backward (or the death of the graph)."""

    def __init__(self, ws):
        self.ws = ws
        ws.busy = True

    def release(self):
        if self.ws is not None:
            self.ws.busy = False
            self.ws = None

    def __del__(self):
        self.release()


class ViTEngine:
    def __init__(self, modules):
        self.modules = list(modules)
        self.G = len(self.modules)
        if self.G not in (1, 2):
            raise MfvError("ViTEngine drives 1 or 2 branches")
        m0 = self.modules[0]
        self.layout = ViTLayout(m0.img_size, m0.patch_size, m0.embed_dim, m0.depth, m0.num_heads, m0.hidden_dim)
        for m in self.modules[1:]:
            if (m.img_size, m.patch_size, m.embed_dim, m.depth, m.num_heads, m.hidden_dim) != \
                    (m0.img_size, m0.patch_size, m0.embed_dim, m0.depth, m0.num_heads, m0.hidden_dim):
                raise MfvError("grouped branches must share one architecture")
        self.device = None
        self.master = self.shadow = self.shadow16 = None
        self.fwd_f16 = FWD_PRECISION == "fp16"
        self.grads = [None, None]
        self.grad_idx = 0
        self._next_grad = None
        self.shadow_fresh = False
        self._shadow_sig = None  # parameter versions the 16-bit shadows were last known to match
        self._casts = 0
        self._ws = {}
        self._params = None  # per group: list of (name, Parameter)
        ref = weakref.ref(self)

        def _on_load(module, incompatible_keys):  # load_state_dict rewrote the fp32 master in place: shadows are stale
            eng = ref()
            if eng is not None:
                eng.invalidate_shadow()

        for m in self.modules:
            m.register_load_state_dict_post_hook(_on_load)

    # ------------------------------------------------------------------------------------------ parameters
    def _named_params(self, g):
        mod = self.modules[g]
        return [(n, mod.get_parameter(n)) for n, _, _ in self.layout.entries]

    def adopt(self, device):
        """Make every encoder Parameter a view into master[g]; re-done transparently if the module was moved/reloaded."""
        lay = self.layout
        if self.master is None or self.device != device:
            self.device = device
            self.master = torch.zeros(self.G, lay.P, device=device, dtype=torch.float32)
            self.shadow = torch.zeros(self.G, lay.P, device=device, dtype=torch.bfloat16)
            self.shadow16 = torch.zeros(self.G, lay.P, device=device, dtype=torch.float16)
            self.grads = [torch.zeros(self.G, lay.P, device=device, dtype=torch.float32), None]
            self._ws = {}
        base = self.master.data_ptr()
        self._params = []
        for g in range(self.G):
            plist = self._named_params(g)
            self._params.append(plist)
            for name, p in plist:
                off = lay.offset[name]
                want = base + 4 * (g * lay.P + off)
                if p.data_ptr() != want or p.device != device:
                    if tuple(p.shape) != lay.shape[name]:
                        raise MfvError("parameter %s has shape %s, expected %s" % (name, tuple(p.shape), lay.shape[name]))
                    view = self.master[g, off:off + p.numel()].view(p.shape)
                    view.copy_(p.detach().to(device=device, dtype=torch.float32))
                    p.data = view
                    self.invalidate_shadow()

    def is_adopted(self):
        if self.master is None or self._params is None:
            return False
        base = self.master.data_ptr()
        lay = self.layout
        for g in range(self.G):
            for name, p in self._params[g]:
                if p.data_ptr() != base + 4 * (g * lay.P + lay.offset[name]):
                    return False
        return True

    def params_flat(self):
        """All encoder Parameters, group-major, layout order (the autograd inputs of the encode Function)."""
        return [p for g in range(self.G) for _, p in self._params[g]]

    def any_requires_grad(self):
        return any(p.requires_grad for g in range(self.G) for _, p in self._params[g])

    # The 16-bit GEMM shadows are valid for ONE forward after somebody who rewrote them said so (mark_shadow_fresh: the
    # fused optimizer step, capture_graph), and only while no Parameter was edited in place since: load_state_dict, a
    # torch.optim step or p.copy_() bump Parameter._version without changing data_ptr.  Writers that bypass autograd's
    # version counter (p.data.copy_, raw kernels such as the EMA update) must call invalidate_shadow().
    def _param_sig(self):
        if self._params is None:
            return None
        # runs on the host in front of every graph replay: a flat cached list and a C-level map instead of a nested
        # generator (300 Parameters: ~30 -> ~10 us that the GPU otherwise idles through when the loop reads the loss back)
        flat = getattr(self, "_params_flat", None)
        if flat is None or self._params_flat_src is not self._params:
            flat = self._params_flat = [p for plist in self._params for _, p in plist]
            self._params_flat_src = self._params
        return sum(map(_VERSION_OF, flat))

    def mark_shadow_fresh(self):
        self.shadow_fresh = True
        self._shadow_sig = self._param_sig()

    def invalidate_shadow(self):
        self.shadow_fresh = False
        self._shadow_sig = None

    def shadow_is_current(self):
        return self.shadow_fresh and self._shadow_sig is not None and self._shadow_sig == self._param_sig()

    def cast_shadow(self):
        ops.cast_shadow(self.master.view(-1), self.shadow.view(-1), self.shadow16.view(-1) if self.fwd_f16 else None)

    def refresh_shadow(self):
        if not self.shadow_is_current():
            self._casts += 1
            if self.fwd_f16 and (self._casts == 1 or self._casts % FP16_CHECK_EVERY == 0):
                self.check_fp16_range()
            self.cast_shadow()
        self.shadow_fresh = False  # a fresh flag is consumed by exactly one forward

    # ------------------------------------------------------------------------------------------ fp16 range policy
    # The fp16-forward mode stores xn, qkv, attn_o and gelu(u) as IEEE half (max 65504).  Instead of testing every
    # element in the issue-bound GEMM epilogues, the engine PROVES the forward cannot overflow from the weights alone:
    # LayerNorm output rows have ||xn||_2 <= sqrt(C) max|gamma| + ||beta||_2, so every output of the Linear that follows
    # is bounded by that times the largest weight-row norm plus the largest bias (Cauchy-Schwarz); attention outputs are
    # convex combinations of v rows; gelu(u) <= |u|.  While the bound stays below FP16_SAFE_BOUND the fp16 forward is
    # safe for ANY input; when it does not (weights of pathological scale), the engine switches itself to the all-bf16
    # forward (same kernels, bf16 operands: no overflow below 3e38, logits within the bf16 tolerance) and warns once.
    def fp16_range_bound(self):
        """Upper bound (device scalar) of |value| over every fp16-stored forward activation, for any input."""
        lay, m = self.layout, self.master
        C_, D = lay.C, lay.depth
        rel = {nm: lay.offset["blocks.0." + nm] - lay.off_block0 for nm, _ in _BLOCK_FIELDS}

        def blk(name, *shape):  # [G, depth, *shape] strided view of one tensor of every block
            st = [lay.P, lay.block_stride]
            acc = 1
            tail = []
            for d in reversed(shape):
                tail.append(acc)
                acc *= d
            return m.as_strided((self.G, D) + tuple(shape), tuple(st + tail[::-1]), lay.off_block0 + rel[name])

        worst = m.new_zeros(())
        for ln, w, b, N in (("norm1", "attn.qkv", "attn.qkv", 3 * C_), ("norm2", "mlp.fc1", "mlp.fc1", lay.hidden)):
            gam, bet = blk(ln + ".weight", C_), blk(ln + ".bias", C_)
            xn_l2 = (C_ ** 0.5) * gam.abs().amax(-1) + bet.norm(dim=-1)            # [G, depth]
            xn_abs = (C_ ** 0.5) * gam.abs().amax(-1) + bet.abs().amax(-1)
            rows = blk(w + ".weight", N, C_).norm(dim=-1).amax(-1)                 # largest weight-row norm
            out = xn_l2 * rows + blk(b + ".bias", N).abs().amax(-1)
            worst = torch.maximum(worst, torch.maximum(out.max(), xn_abs.max()))
        return worst

    def check_fp16_range(self):
        """Switches the engine to the bf16 forward when the fp16 one is not provably overflow-free.  Returns the bound."""
        if not self.fwd_f16 or self.master is None:
            return None
        bound = float(self.fp16_range_bound())
        if not (bound < FP16_SAFE_BOUND):  # also catches NaN / inf weights
            import warnings
            warnings.warn("mfvit: forward activations may reach %.3g, beyond the fp16 range bound %.3g for these weights; "
                          "this encoder now runs its forward with bf16 operands (MFVIT_FWD_PRECISION=bf16 selects that "
                          "from the start)" % (bound, FP16_SAFE_BOUND))
            self.fwd_f16 = False
            self.invalidate_shadow()
        return bound

    # ------------------------------------------------------------------------------------------ plan / workspaces
    def _workspace(self, B, save):
        lst = self._ws.setdefault((B, save, self.fwd_f16), [])
        for ws in lst:
            if not ws.busy:
                return ws
        ws = _Workspace(self.layout, self.G, B, save, self.device, self.fwd_f16)
        lst.append(ws)
        return ws

    def _plan(self, ws, images, tokens, dtokens=None, grad=None, stop_grad_conv1=False):
        lay = self.layout
        p = VitPlan()
        p.G, p.B, p.S, p.C, p.H, p.depth, p.hidden, p.img, p.np, p.P = (
            self.G, ws.B, lay.S, lay.C, lay.H, lay.depth, lay.hidden, lay.img, lay.np, lay.P)
        p.master, p.shadow, p.shadow16 = self.master.data_ptr(), self.shadow.data_ptr(), self.shadow16.data_ptr()
        p.fwd_f16 = 1 if self.fwd_f16 else 0
        p.gact_bf_per_block = 1 if getattr(ws, "gact_twin", False) else 0
        for f in ("patches_bf", "xn_bf", "attn_o_bf", "gact_bf"):
            t = getattr(ws, f)
            setattr(p, f, t.data_ptr() if t is not None else None)
        p.grad = grad.data_ptr() if grad is not None else None
        o = lay.offset
        p.off_cls, p.off_pos, p.off_pe_w, p.off_pe_b = (o["cls_token"], o["pos_embed"], o["patch_embed.proj.weight"],
                                                        o["patch_embed.proj.bias"])
        p.off_norm_w, p.off_norm_b = o["norm.weight"], o["norm.bias"]
        p.off_block0, p.block_stride = lay.off_block0, lay.block_stride
        for nm, field in _BLOCK_FIELDS:
            setattr(p, field, o["blocks.0." + nm] - lay.off_block0)
        for g in range(self.G):
            p.images[g] = images[g].data_ptr() if images is not None else None
        for f in ("patches", "acc", "x", "xn", "stats", "qkv", "attn_o", "lse", "u", "gact"):
            setattr(p, f, getattr(ws, f).data_ptr())
        p.tokens = tokens.data_ptr() if tokens is not None else None
        p.save_for_backward = 1 if ws.save else 0
        p.stop_grad_conv1 = 1 if stop_grad_conv1 else 0
        if dtokens is not None:
            b = ws.ensure_bwd(lay, self.device)
            p.dtokens = dtokens.data_ptr()
            p.dx[0], p.dx[1] = b["dx0"].data_ptr(), b["dx1"].data_ptr()
            p.dx16[0], p.dx16[1] = b["dx16_0"].data_ptr(), b["dx16_1"].data_ptr()
            for f in ("dhid", "dxn", "d_o", "dqkv", "delta", "dacc"):
                setattr(p, f, b[f].data_ptr())
            p.attn_ws = b["attn_ws"].data_ptr() if b["attn_ws"] is not None else None
        return p

    # ------------------------------------------------------------------------------------------ forward / backward
    def forward(self, images, save):
        """images: list of G fp32 [B,3,H,W] CUDA tensors -> tokens f32 [G,B,S,C] (+ a lease when save=True)."""
        dev = images[0].device
        if dev.type != "cuda":
            raise MfvError("the MF-ViT encoder runs only on CUDA sm_100a devices; got %s (no CPU fallback)" % dev)
        lib = _lib.init(dev.index if dev.index is not None else torch.cuda.current_device())
        if not self.is_adopted() or self.device != dev:
            self.adopt(dev)
        lay = self.layout
        B = images[0].shape[0]
        imgs = []
        for im in images:
            if tuple(im.shape) != (B, 3, lay.img, lay.img):
                raise MfvError("expected images of shape %s, got %s" % ((B, 3, lay.img, lay.img), tuple(im.shape)))
            imgs.append(im.detach().to(torch.float32).contiguous())
        self.refresh_shadow()
        ws = self._workspace(B, bool(save))
        tokens = torch.empty(self.G, B, lay.S, lay.C, device=dev, dtype=torch.float32)
        plan = self._plan(ws, imgs, tokens)
        check(lib.mfv_vit_forward(C.byref(plan), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
              "mfv_vit_forward")
        lease = _Lease(ws) if save else None
        ws.images = imgs if save else None  # the backward rebuilds the bf16 patch matrix from them (conv weight gradient)
        self._last_images = imgs  # keep inputs alive until the stream has consumed them
        return tokens, lease

    def _pick_grad_buffer(self):
        """Ping-pong so that a view autograd kept from the previous backward is never overwritten in place."""
        lay = self.layout
        nxt = self.grad_idx ^ 1
        if self.grads[nxt] is None:
            self.grads[nxt] = torch.zeros(self.G, lay.P, device=self.device, dtype=torch.float32)
        lo = self.grads[nxt].data_ptr()
        hi = lo + 4 * self.G * lay.P
        for g in range(self.G):
            for _, p in self._params[g]:
                if p.grad is not None and lo <= p.grad.data_ptr() < hi:
                    self.grads[nxt] = torch.zeros(self.G, lay.P, device=self.device, dtype=torch.float32)
                    break
        self.grad_idx = nxt
        return self.grads[nxt]

    def next_grad_buffer(self):
        """The flat buffer the NEXT backward will write (picked now so that a caller can zero it ahead of time, off the
        critical path, and pass zero=False to backward)."""
        if self._next_grad is None:
            self._next_grad = self._pick_grad_buffer()
        return self._next_grad

    def backward(self, lease, dtokens, zero=True, segments=None, on_segment=None):
        """Encoder backward.  `segments` (optional): list of (block_hi, block_lo) covering depth-1 .. 0 in descending
        order; `on_segment(grad, lo_elem, hi_elem)` is called after each one with the element range of every group's flat
        gradient buffer that is now final (data-parallel callers start that slice's all-reduce there)."""
        if lease is None or lease.ws is None:
            raise MfvError("backward called without saved activations (forward ran with save=False?)")
        ws = lease.ws
        lib = _lib.init(self.device.index)
        grad = self._next_grad if self._next_grad is not None else self._pick_grad_buffer()
        self._next_grad = None
        if zero:
            ops.fill_(grad.view(-1), 0.0)
        dtokens = dtokens.contiguous()
        stop = not self._params[0][2][1].requires_grad  # patch_embed.proj.weight frozen (stop_grad_conv1)
        plan = self._plan(ws, getattr(ws, "images", None), None, dtokens=dtokens, grad=grad, stop_grad_conv1=stop)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        if not segments:
            check(lib.mfv_vit_backward(C.byref(plan), stream), "mfv_vit_backward")
        else:
            lay = self.layout
            upper = lay.P
            for i, (hi, lo) in enumerate(segments):
                flags = (1 if i == 0 else 0) | (2 if i == len(segments) - 1 else 0)
                check(lib.mfv_vit_backward_range(C.byref(plan), stream, int(hi), int(lo), flags),
                      "mfv_vit_backward_range")
                lower = 0 if i == len(segments) - 1 else lay.off_block0 + lo * lay.block_stride
                if on_segment is not None:
                    on_segment(grad, lower, upper)
                upper = lower
        lease.release()
        return grad

    def grad_views(self, grad):
        lay = self.layout
        out = []
        for g in range(self.G):
            for name, p in self._params[g]:
                if p.requires_grad:
                    off = lay.offset[name]
                    out.append(grad[g, off:off + p.numel()].view(p.shape))
                else:
                    out.append(None)
        return out


class _EncodeFn(torch.autograd.Function):
    """tokens = encoder(images): autograd seam around ViTEngine (parameters are passed so DDP / optimisers see them)."""

    @staticmethod
    def forward(ctx, engine, save, n_img, *tensors):
        images = tensors[:n_img]
        tokens, lease = engine.forward(list(images), save)
        ctx.engine, ctx.lease, ctx.n_img = engine, lease, n_img
        return tokens

    @staticmethod
    def backward(ctx, dtokens):
        eng = ctx.engine
        grad = eng.backward(ctx.lease, dtokens)
        return (None, None, None) + (None,) * ctx.n_img + tuple(eng.grad_views(grad))


def encode(engine, images):
    """Run the grouped encoder with autograd support.  Returns tokens f32 [G,B,S,C]."""
    dev = images[0].device
    if dev.type != "cuda":
        raise MfvError("the MF-ViT encoder runs only on CUDA sm_100a devices; got %s (no CPU fallback)" % dev)
    if not engine.is_adopted() or engine.device != dev:
        engine.adopt(dev)
    need = torch.is_grad_enabled() and engine.any_requires_grad()
    if not need:
        with torch.no_grad():
            tokens, _ = engine.forward(list(images), save=False)
        return tokens
    return _EncodeFn.apply(engine, True, len(images), *images, *engine.params_flat())


_ENGINES = weakref.WeakValueDictionary()


def engine_for(*modules):
    """One cached engine per (ordered) tuple of backbone modules."""
    key = tuple(id(m) for m in modules)
    eng = _ENGINES.get(key)
    if eng is None or any(a is not b for a, b in zip(eng.modules, modules)):
        eng = ViTEngine(modules)
        _ENGINES[key] = eng
        for m in modules:
            m._mfv_engines = getattr(m, "_mfv_engines", [])
            m._mfv_engines.append(eng)  # keep alive with the module
    return eng
