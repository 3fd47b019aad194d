"""DeepLabV3+ over a ResNet encoder, forward AND backward, on the CUDA layers of libaadg_b200.so.

Keeps the constructor and call shape the reference uses (models/__init__.py:17-23, models/heads.py:14-25):

    model = DeepLabV3Plus(encoder_name="resnet50", encoder_weights=None, in_channels=3, classes=2,
                          aux_params=dict(pooling="avg"))
    logits, pooled = model(x)          # x float32 [N,3,H,W] in [-1,1]

and the state_dict key names of segmentation_models_pytorch 0.2.0 (encoder.*, decoder.aspp.*, ...), but
runs on its own engine: bf16 NHWC activations, tcgen05 implicit-GEMM convolutions, fused batch-norm /
ReLU / residual / dropout kernels, one fused up-sample + sigmoid + BCE (+ Dice counts) loss kernel, an
explicit backward pass (no autograd graph) and a fused Adam over one flat parameter buffer.
Architecture: SURVEY.md App. A.3 (smp DeepLabV3Plus, output stride 16, ASPP rates 12/24/36, separable).
"""
import math
import os

import numpy as np
import torch

from ..ops import conv as C
from ..ops import nn as K

BF16 = torch.bfloat16
BN_EPS, BN_MOMENTUM = 1e-5, 0.1
# batch-norm statistics accumulated in the convolution epilogue (aadg_conv_fprop_stats_bf16) instead of a separate pass
FUSE_BN_STATS = os.environ.get("AADG_FUSE_BN_STATS", "1") != "0"
# residual blocks hand their two input-gradient branches to the next batch-norm backward as a pair (added on load)
# instead of accumulating one into the other in the convolution epilogue
GRAD_PAIRS = os.environ.get("AADG_GRAD_PAIRS", "0") != "0"   # measured: 1.7 ms/step slower than accumulating
# ... except for the 1x1 expansion convolutions with few input channels (Cout >= 4*Cin, Cin <= UNFUSE_MAX_CIN): those
# are bound by writing their output, the statistics warps slow that epilogue down by more than a separate read-only pass
# over the output costs (per-layer A/B, profiles/: 64->256 @128^2 1.68 ms fused vs 1.15 + 0.22 ms separate)
UNFUSE_MAX_CIN = int(os.environ.get("AADG_UNFUSE_MAX_CIN", "256"))
# Optional (AADG_WGRAD_SIDE=1): the weight gradient of a layer is off the backward pass's critical path (nothing reads dW
# before the gradient exchange / the optimiser), so it can run on a SIDE stream, concurrently with the data gradient of
# the same layer and the batch-norm backward of the next one; joined back before every gradient-ready hook and at the end
# of the backward pass (works inside CUDA-graph capture: fork / join).  MEASURED NEUTRAL on the 144 x 512^2 ResNet-50 step
# (89.7 vs 90.0 ms, graph and eager, profiles/README.md): the persistent convolution kernels own every SM's shared
# memory, the element-wise passes they would overlap are HBM-bound like the weight gradient's own operand streams, and the
# step has no idle pipe to fill.  Off by default.
WGRAD_SIDE_STREAM = os.environ.get("AADG_WGRAD_SIDE", "0") != "0"


class _Side:
    streams = {}          # device index -> the side stream
    pending = []          # tensors the side stream still reads (kept alive until the join)
    forked = None         # the side stream with un-joined work
    armed = False         # only a whole-network backward (loss_step / the autograd surface) forks: it also joins


def on_side(fn, *keep):
    """run fn() (kernel launches) on the side stream, ordered after everything enqueued so far on the current stream"""
    if not (WGRAD_SIDE_STREAM and _Side.armed) or C.TIMING is not None:      # per-launch event timing wants one stream
        fn()
        return
    dev = torch.cuda.current_device()
    side = _Side.streams.get(dev)
    if side is None:
        side = _Side.streams[dev] = torch.cuda.Stream(device=dev)
    ev = torch.cuda.Event()
    ev.record()
    side.wait_event(ev)
    with torch.cuda.stream(side):
        fn()
    _Side.pending.extend(keep)
    _Side.forked = side


def join_side():
    """the current stream waits for the side stream's work; the tensors it was reading may be released afterwards"""
    side = _Side.forked
    if side is not None:
        ev = torch.cuda.Event()
        ev.record(side)
        torch.cuda.current_stream().wait_event(ev)
        _Side.forked = None
    _Side.pending.clear()


# ---------------------------------------------------------------------------------------------------
# parameters: one flat fp32 buffer (+ grads, Adam moments), bf16 copies of the conv weights
# ---------------------------------------------------------------------------------------------------
class Param:
    __slots__ = ("name", "shape", "kind", "offset", "numel", "init", "bf_off", "bft_off", "store")

    def __init__(self, name, shape, kind, init):
        self.name, self.shape, self.kind, self.init = name, tuple(shape), kind, init
        self.numel = int(np.prod(shape))
        self.offset = self.bf_off = self.bft_off = -1
        self.store = None

    @property
    def data(self):
        return self.store.params[self.offset:self.offset + self.numel].view(self.shape)

    @property
    def grad(self):
        return self.store.grads[self.offset:self.offset + self.numel].view(self.shape)

    @property
    def bf16(self):
        return self.store.wb[self.bf_off:self.bf_off + self.numel].view(self.shape)

    @property
    def bf16_t(self):
        t, co, ci = self.shape
        return self.store.wbt[self.bft_off:self.bft_off + self.numel].view(t, ci, co)


class ParamStore:
    """kinds: 'conv' (bf16 copy + transposed copy), 'conv_nt' (bf16 copy only), 'f32' (used as is)."""

    def __init__(self, device):
        self.device = device
        self.items = []
        self.params = None

    def add(self, name, shape, kind, init):
        assert self.params is None
        p = Param(name, shape, kind, init)
        p.store = self
        self.items.append(p)
        return p

    def finalize(self):
        off = bf = bft = 0
        for p in self.items:
            p.offset = off
            off += (p.numel + 3) // 4 * 4
            if p.kind in ("conv", "conv_nt"):
                p.bf_off = bf
                bf += (p.numel + 7) // 8 * 8
            if p.kind == "conv":
                p.bft_off = bft
                bft += (p.numel + 7) // 8 * 8
        dev = self.device
        self.params = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grads = torch.zeros(off, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(off, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(off, dtype=torch.float32, device=dev)
        self.wb = torch.zeros(max(bf, 8), dtype=BF16, device=dev)
        self.wbt = torch.zeros(max(bft, 8), dtype=BF16, device=dev)
        descs = []
        for p in self.items:
            if p.kind in ("conv", "conv_nt"):
                t, co, ci = p.shape
                descs.append(np.array([p.offset, p.bf_off, p.bft_off if p.kind == "conv" else -1], np.int64).tobytes()
                             + np.array([t, co, ci, 0], np.int32).tobytes())
        self.n_descs = len(descs)
        self.descs = torch.frombuffer(bytearray(b"".join(descs)), dtype=torch.uint8).to(dev)
        for p in self.items:
            p.data.copy_(p.init(p.shape).to(dev))
            p.init = None
        self.step = 0
        # the optimiser's scalars in device memory (aadg_adam_step_dev): nothing is passed by value, so a captured CUDA
        # graph replays the update while the host changes the rate / the step count advances on the device
        self.hyper = torch.tensor([1e-3, 0.9, 0.999, 1e-8, 0.0, 1.0], dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.refresh()

    def refresh(self):
        """bf16 (and transposed) copies of the convolution weights from the fp32 masters."""
        K.weight_prep(self.params, self.wb, self.wbt, self.descs, self.n_descs)
        self._synced_version = self.params._version

    def sync(self):
        """refresh the bf16 copies if a torch in-place op (an optimizer stepping the parameter leaves, copy_) touched the
        masters since the last refresh; the engine's own Adam kernels refresh by themselves"""
        if self.params._version != getattr(self, "_synced_version", -1):
            self.refresh()

    def zero_grad(self):
        self.grads.zero_()

    def adam_step(self, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.step += 1
        self.step_dev += 1
        K.adam_step(self.params, self.grads, self.exp_avg, self.exp_avg_sq, lr, betas[0], betas[1], eps,
                    weight_decay, self.step)
        self.refresh()

    def set_hyper(self, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0):
        """(host -> device copy; call it outside a captured region) grad_scale multiplies the gradients on load:
        1/world after a summing all-reduce"""
        self.hyper.copy_(torch.tensor([lr, betas[0], betas[1], eps, weight_decay, grad_scale], dtype=torch.float32))

    def adam_step_dev(self):
        """the same Adam step driven by `hyper` / `step_dev` in device memory (capturable)"""
        self.step += 1
        self.step_dev += 1
        K.adam_step_dev(self.params, self.grads, self.exp_avg, self.exp_avg_sq, self.hyper, self.step_dev)
        self.refresh()


def kaiming_fan_out(shape_torch):
    """torchvision ResNet init: kaiming_normal_(mode='fan_out', nonlinearity='relu') on [Cout,Cin,R,S]."""
    co, ci, r, s = shape_torch
    return torch.randn(shape_torch) * math.sqrt(2.0 / (co * r * s))


def kaiming_uniform_default(shape_torch):
    """nn.Conv2d default init (kaiming_uniform_(a=sqrt(5))) used by smp's decoder convs."""
    co, ci, r, s = shape_torch
    bound = 1.0 / math.sqrt(ci * r * s)
    return (torch.rand(shape_torch) * 2 - 1) * bound


def to_taps(w):
    """torch [Cout,Cin,R,S] -> [R*S,Cout,Cin]"""
    co, ci, r, s = w.shape
    return w.permute(2, 3, 0, 1).reshape(r * s, co, ci).contiguous()


def from_taps(w, r, s):
    t, co, ci = w.shape
    return w.reshape(r, s, co, ci).permute(2, 3, 0, 1).contiguous()


# ---------------------------------------------------------------------------------------------------
# layers
# ---------------------------------------------------------------------------------------------------
# SyncBN option: a torch.distributed process group (or True for the default group) makes every
# BatchNorm share its batch statistics across ranks -- forward sums and the two backward sums are all-reduced -- so that
# N ranks holding shards of a batch compute what one GPU holding the whole batch computes.  None = per-rank statistics,
# the reference's DDP behaviour (plain BatchNorm2d under DistributedDataParallel, models/__init__.py:39).
SYNC_BN = None


def set_sync_bn(group):
    global SYNC_BN
    SYNC_BN = group


def _sync_world():
    import torch.distributed as dist
    if SYNC_BN is None or not (dist.is_available() and dist.is_initialized()):
        return None, 1
    g = None if SYNC_BN is True else SYNC_BN
    return g, dist.get_world_size(g)


def _all_reduce_sum(t, group):
    import torch.distributed as dist
    dist.all_reduce(t, group=group)


class BatchNorm:
    def __init__(self, store, name, c):
        self.c, self.name = c, name
        self.gamma = store.add(name + ".weight", (c,), "f32", lambda s: torch.ones(s))
        self.beta = store.add(name + ".bias", (c,), "f32", lambda s: torch.zeros(s))
        dev = store.device
        self.running_mean = torch.zeros(c, device=dev)
        self.running_var = torch.ones(c, device=dev)
        self.num_batches_tracked = 0
        self.buf = torch.zeros(6, c, device=dev)    # sum, sumsq, mean, invstd, scale, shift

    def stats_buffers(self):
        """zeroed (sum, sumsq) accumulators for a convolution epilogue to fill (C.fprop(stats=...)): they start at zero
        and bn_finalize(reset_sums=True) zeroes them again after reading"""
        return self.buf[0], self.buf[1]

    def forward(self, x, y, training, res=None, relu=True, dropout_seed=None, relu_bits=None, have_stats=False,
                relu6=False):
        s = self.buf
        if training:
            if not have_stats:
                s[:2].zero_()
                K.bn_stats(x, s[0], s[1])
            count = x.numel() // x.shape[-1]
            group, world = _sync_world()
            if world > 1:
                _all_reduce_sum(s[:2], group)
                count *= world
            self.count = count
            K.bn_finalize(s[0], s[1], self.gamma.data, self.beta.data, count, BN_EPS, BN_MOMENTUM, s[2], s[3], s[4],
                          s[5], self.running_mean, self.running_var, reset_sums=True)
            self.num_batches_tracked += 1
        else:
            torch.rsqrt(self.running_var + BN_EPS, out=s[3])
            s[2].copy_(self.running_mean)
            torch.mul(self.gamma.data, s[3], out=s[4])
            torch.sub(self.beta.data, s[2] * s[4], out=s[5])
        K.bn_apply(x, s[4], s[5], y, res=res, relu=relu, dropout_seed=dropout_seed, relu_bits=relu_bits, relu6=relu6)
        self.saved = s[2:6] if training else None       # mean, invstd, scale, shift (this layer's own buffer)

    def backward(self, dy, x, y, dx, relu=True, dropout_seed=None, dres=None, dres_accumulate=False, dy2=None,
                 relu6=False):
        """y=None: no residual was added in forward, the ReLU mask is recomputed from x (saves reading y).
        dy2: second gradient branch, added to dy on load."""
        sv = self.saved
        group, world = _sync_world()
        if world > 1:
            # local sums -> this rank's share of dgamma / dbeta (the gradient all-reduce sums the shares); the GLOBAL
            # sums and the global pixel count give dx
            sums = torch.zeros((2, self.c), dtype=torch.float32, device=x.device)
            K.bn_backward_reduce(dy, x, y, sv[0], sv[1], self.gamma.data, sums, relu=relu, dropout_seed=dropout_seed,
                                 shift=sv[3], dy2=dy2, relu6=relu6)
            self.gamma.grad.add_(sums[0])
            self.beta.grad.add_(sums[1])
            _all_reduce_sum(sums, group)
            K.bn_backward_apply(dy, x, y, sv[0], sv[1], self.gamma.data, sums, self.count, dx, relu=relu,
                                dropout_seed=dropout_seed, dres=dres, dres_accumulate=dres_accumulate, shift=sv[3],
                                dy2=dy2, relu6=relu6)
            self.saved = None
            return
        K.bn_backward(dy, x, y, sv[0], sv[1], self.gamma.data, self.gamma.grad, self.beta.grad, dx, relu=relu,
                      dropout_seed=dropout_seed, dres=dres, dres_accumulate=dres_accumulate, shift=sv[3], dy2=dy2,
                      grads_zeroed=True,       # gamma.grad / beta.grad: slices of the flat buffer zero_grad() cleared
                      relu6=relu6)
        self.saved = None


# ---------------------------------------------------------------------------------------------------
# pixel packing for 16 / 32-channel 3x3 convolutions on wide maps
# ---------------------------------------------------------------------------------------------------
# A [N,H,W,16] tensor has 32-byte pixel rows: the TMA moves one short row per pixel and the tensor cores see K = 16.
# Viewing P = 64 / Cin consecutive pixels as ONE 64-"channel" pixel ([N,H,W/P,64], the same memory) turns the layer
# into a 64 -> P*Cout 3x3 convolution on a W/P-wide map whose weights are the original taps scattered into P x P
# blocks: macro tap (dy, T) holds w[dy, dx] at block (b, a) iff dx = P*T + a - b is a real tap (-1, 0, 1).  Full
# 128-byte rows, the halo-reuse kernels apply, and the 16x arithmetic inflation is free on idle tensor cores.
PACK_MIN_W = int(os.environ.get("AADG_PACK_MIN_W", "65"))    # packed width from which the halo kernels pay off


class PackedTaps:
    """index maps between w [9, Co, Ci] and its packed expansion W4 [9, P*Co, P*Ci] (built once per layer)"""

    def __init__(self, cout, cin, P, device):
        src = np.full((3, 3, P, cout, P, cin), -1, np.int64)          # [dy, T, b, co, a, ci] -> flat index into w
        base = np.arange(9 * cout * cin, dtype=np.int64).reshape(3, 3, cout, cin)   # [dy, dx, co, ci]
        for T in (-1, 0, 1):
            for a in range(P):
                for b in range(P):
                    dx = P * T + a - b
                    if -1 <= dx <= 1:
                        src[:, T + 1, b, :, a, :] = base[:, dx + 1]
        src = src.reshape(-1)
        self.P, self.cout, self.cin = P, cout, cin
        self.shape = (9, P * cout, P * cin)
        valid = np.nonzero(src >= 0)[0]
        self.valid = torch.from_numpy(valid).to(device)                # positions of W4 that hold a real tap
        self.src = torch.from_numpy(src[valid]).to(device)             # ... and the w element each one copies
        # transposed expansion for the data gradient: [9, P*Ci, P*Co]
        perm = np.arange(src.size, dtype=np.int64).reshape(9, P * cout, P * cin).transpose(0, 2, 1).reshape(-1)
        inv = np.empty_like(perm)
        inv[perm] = np.arange(perm.size)
        self.valid_t = torch.from_numpy(inv[valid]).to(device)

    def expand(self, w_bf16):
        out = torch.zeros(self.shape, dtype=BF16, device=w_bf16.device)
        out.view(-1)[self.valid] = w_bf16.reshape(-1)[self.src]
        return out

    def expand_t(self, w_bf16):
        out = torch.zeros((9, self.shape[2], self.shape[1]), dtype=BF16, device=w_bf16.device)
        out.view(-1)[self.valid_t] = w_bf16.reshape(-1)[self.src]
        return out

    def fold_grad(self, dw4, dw):
        """dw [9, Co, Ci] (fp32, accumulated) += the real taps scattered over dW4"""
        dw.view(-1).index_add_(0, self.src, dw4.reshape(-1)[self.valid])


class ConvBN:
    """Conv2d(bias=False) -> BatchNorm2d -> [+ residual] -> [ReLU] -> [Dropout(0.5)]"""

    def __init__(self, store, conv_name, bn_name, cin, cout, k=1, stride=1, pad=0, dil=1, relu=True, init=None,
                 need_dgrad=True):
        self.cin, self.cout, self.k, self.stride, self.pad, self.dil = cin, cout, k, stride, pad, dil
        self.relu, self.relu6 = bool(relu), relu == "relu6"
        init = init or kaiming_fan_out
        self.w = store.add(conv_name + ".weight", (k * k, cout, cin), "conv" if need_dgrad else "conv_nt",
                           lambda s: to_taps(init((cout, cin, k, k))))
        self.bn = BatchNorm(store, bn_name, cout)
        self.need_dgrad = need_dgrad
        self.packed = None        # PackedTaps when the layer qualifies (decided at the first forward)

    def out_hw(self, h, w):
        return C.out_size(h, w, self.k, self.k, self.stride, self.pad, self.dil)

    def _pack_factor(self, x):
        """P > 1 when this call runs pixel-packed: 3x3 / stride 1 / pad 1 / dilation 1, 16 or 32 dense input channels"""
        if (self.k, self.stride, self.pad, self.dil) != (3, 1, 1, 1) or self.cin not in (16, 32):
            return 0
        P = 64 // self.cin
        w = x.shape[2]
        if x.stride(2) != self.cin or not x.is_contiguous() or w % P or w // P < PACK_MIN_W or P * self.cout > 64:
            return 0
        return P

    def _fprop_packed(self, x, P, fused):
        n, h, w, _ = x.shape
        if self.packed is None:
            self.packed = PackedTaps(self.cout, self.cin, P, x.device)
        pre = torch.empty((n, h, w, self.cout), dtype=BF16, device=x.device)
        stats = None
        if fused:
            stats = (torch.zeros(P * self.cout, device=x.device), torch.zeros(P * self.cout, device=x.device))
        C.fprop(x.view(n, h, w // P, 64), self.packed.expand(self.w.bf16), 3, 3, 1, 1, 1,
                out=pre.view(n, h, w // P, P * self.cout), stats=stats,
                flops=2.0 * n * h * w * self.cout * self.cin * 9)      # the layer's own FLOPs, not the packed problem's
        if fused:      # packed channel (b, co) -> co
            s0, s1 = self.bn.stats_buffers()
            s0.add_(stats[0].view(P, self.cout).sum(0))
            s1.add_(stats[1].view(P, self.cout).sum(0))
        return pre

    def forward(self, x, training, out=None, res=None, dropout_seed=None):
        n, h, w, _ = x.shape
        ho, wo = self.out_hw(h, w)
        fused = training and FUSE_BN_STATS and not (self.k == 1 and self.cout >= 4 * self.cin and
                                                    self.cin <= UNFUSE_MAX_CIN)
        P = self._pack_factor(x)
        if P:
            pre = self._fprop_packed(x, P, fused)
        else:
            pre = C.fprop(x, self.w.bf16, self.k, self.k, self.stride, self.pad, self.dil,
                          stats=self.bn.stats_buffers() if fused else None)
        if out is None:
            out = torch.empty((n, ho, wo, self.cout), dtype=BF16, device=x.device)
        # backward needs the ReLU mask: recomputed from `pre` when there is no residual, otherwise kept as one
        # bit per element (reading it costs 1/16 of re-reading the output tensor)
        bits = None
        if training and res is not None and self.relu:
            bits = torch.empty((pre.numel() // 8,), dtype=torch.uint8, device=x.device)
        self.bn.forward(pre, out, training, res=res, relu=self.relu, dropout_seed=dropout_seed, relu_bits=bits,
                        have_stats=fused, relu6=self.relu6)
        self.ctx = (x, pre, bits, dropout_seed, P) if training else None
        return out

    def backward(self, dy, dx=None, accumulate=False, want_dres=False, dres=None, dres_accumulate=False, dy2=None):
        """returns dx (or None when the input needs no gradient); with want_dres also the residual gradient.
        The incoming gradient is dy (+ dy2 when given)."""
        x, pre, y, seed, P = self.ctx
        self.ctx = None
        dpre = torch.empty_like(pre)
        if want_dres and dres is None:
            dres = torch.empty(pre.shape, dtype=BF16, device=pre.device)
        self.bn.backward(dy, pre, y, dpre, relu=self.relu, dropout_seed=seed, dres=dres if want_dres else None,
                         dres_accumulate=dres_accumulate, dy2=dy2, relu6=self.relu6)
        if P and (dx is None or (dx.stride(2) == self.cin and not accumulate)):
            n, h, w, _ = x.shape
            x4, d4 = x.view(n, h, w // P, 64), dpre.view(n, h, w // P, P * self.cout)
            real_flops = 2.0 * n * h * w * self.cout * self.cin * 9
            def packed_wgrad():
                dw4 = C.wgrad(x4, d4, 3, 3, 1, 1, 1, flops=real_flops)
                self.packed.fold_grad(dw4, self.w.grad)
                _Side.pending.append(dw4)
            on_side(packed_wgrad, x, dpre)
            if self.need_dgrad:
                if dx is None:
                    dx = torch.empty(x.shape, dtype=BF16, device=x.device)
                C.dgrad(d4, self.packed.expand_t(self.w.bf16), 3, 3, 1, 1, 1, (h, w // P), out=dx.view(n, h, w // P, 64),
                        flops=real_flops)
            else:
                dx = None
            return (dx, dres) if want_dres else dx
        on_side(lambda: C.wgrad(x, dpre, self.k, self.k, self.stride, self.pad, self.dil, out=self.w.grad), x, dpre)
        if self.need_dgrad:
            dx = C.dgrad(dpre, self.w.bf16_t, self.k, self.k, self.stride, self.pad, self.dil, x.shape[1:3], out=dx,
                         accumulate=accumulate)
        else:
            dx = None
        return (dx, dres) if want_dres else dx


class Depthwise3x3:
    def __init__(self, store, name, c, dil, stride=1, init=None):
        self.c, self.dil, self.stride = c, dil, stride
        init = init or kaiming_uniform_default
        self.w = store.add(name + ".weight", (9, c), "f32",
                           lambda s: init((c, 1, 3, 3)).reshape(c, 9).t().contiguous())

    def forward(self, x, training):
        n, h, w, c = x.shape
        ho, wo = (h - 1) // self.stride + 1, (w - 1) // self.stride + 1
        y = torch.empty((n, ho, wo, c), dtype=BF16, device=x.device)
        K.dwconv3x3(x, self.w.data, self.dil, y, stride=self.stride)
        self.ctx = x if training else None
        return y

    def backward(self, dy, dx=None, accumulate=False):
        """dx given with accumulate: the data gradient is added to it (fan-in of parallel branches, no separate add)"""
        x = self.ctx
        self.ctx = None
        on_side(lambda: K.dwconv3x3_wgrad(x, dy, self.dil, self.w.grad, stride=self.stride), x, dy)
        if dx is None:
            dx, accumulate = torch.empty(x.shape, dtype=BF16, device=x.device), False
        K.dwconv3x3(dy, self.w.data, self.dil, dx, backward_data=True, stride=self.stride, accumulate=accumulate)
        return dx


def first_offset(obj):
    """smallest flat-buffer offset of the parameters under `obj` (layers register their parameters in forward order, so
    everything from this offset to the next module's first offset belongs to `obj`)"""
    best = [None]

    def walk(o):
        if isinstance(o, Param):
            best[0] = o.offset if best[0] is None else min(best[0], o.offset)
        elif isinstance(o, (list, tuple)):
            for i in o:
                walk(i)
        elif hasattr(o, "__dict__") and not isinstance(o, (ParamStore, torch.Tensor)):
            for v in vars(o).values():
                walk(v)
    walk(obj)
    return best[0]


def as_pair(d):
    return d if isinstance(d, tuple) else (d, None)


def fold_pair(d):
    """materialise a gradient pair: a += b"""
    a, b = as_pair(d)
    if b is not None:
        K.add_(a, b)
    return a


def _identity_grad(ds, dres, extra):
    """gradient of a residual block's input through its identity / downsample branch, plus an optional extra input
    gradient: the downsample convolution's data gradient accumulates straight into `extra`"""
    if ds is not None:
        if extra is not None:
            return ds.backward(dres, dx=extra, accumulate=True)
        return ds.backward(dres)
    if extra is not None:
        K.add_(dres, extra)
    return dres


class Bottleneck:
    expansion = 4

    def __init__(self, store, name, cin, planes, stride, dil, downsample):
        self.c1 = ConvBN(store, name + ".conv1", name + ".bn1", cin, planes, 1)
        self.c2 = ConvBN(store, name + ".conv2", name + ".bn2", planes, planes, 3, stride, dil, dil)
        self.c3 = ConvBN(store, name + ".conv3", name + ".bn3", planes, planes * 4, 1)
        self.ds = ConvBN(store, name + ".downsample.0", name + ".downsample.1", cin, planes * 4, 1, stride,
                         relu=False) if downsample else None

    def forward(self, x, training):
        idt = self.ds.forward(x, training) if self.ds else x
        return self.c3.forward(self.c2.forward(self.c1.forward(x, training), training), training, res=idt)

    def _identity_grad(self, dres, extra):
        return _identity_grad(self.ds, dres, extra)

    def backward(self, dy, extra=None):
        """dy: tensor or pair of tensors whose sum is the gradient of the block output; `extra`: another gradient of
        the block INPUT (a decoder skip) that is merged here instead of by a separate add pass.  Returns the gradient
        of the block input (a pair (identity / downsample branch, conv branch) with AADG_GRAD_PAIRS=1)."""
        dy, dy2 = as_pair(dy)
        d2, dres = self.c3.backward(dy, want_dres=True, dy2=dy2)
        d1 = self.c2.backward(d2)
        if not GRAD_PAIRS:      # the conv branch accumulates into the identity branch (TMA reduce-add epilogue)
            return self.c1.backward(d1, dx=self._identity_grad(dres, extra), accumulate=True)
        return self._identity_grad(dres, extra), self.c1.backward(d1)


class BasicBlock:
    expansion = 1

    def __init__(self, store, name, cin, planes, stride, dil, downsample):
        self.c1 = ConvBN(store, name + ".conv1", name + ".bn1", cin, planes, 3, stride, dil, dil)
        self.c2 = ConvBN(store, name + ".conv2", name + ".bn2", planes, planes, 3, 1, dil, dil)
        self.ds = ConvBN(store, name + ".downsample.0", name + ".downsample.1", cin, planes, 1, stride,
                         relu=False) if downsample else None

    def forward(self, x, training):
        idt = self.ds.forward(x, training) if self.ds else x
        return self.c2.forward(self.c1.forward(x, training), training, res=idt)

    def _identity_grad(self, dres, extra):
        return _identity_grad(self.ds, dres, extra)

    def backward(self, dy, extra=None):
        """see Bottleneck.backward"""
        dy, dy2 = as_pair(dy)
        d1, dres = self.c2.backward(dy, want_dres=True, dy2=dy2)
        if not GRAD_PAIRS:
            return self.c1.backward(d1, dx=self._identity_grad(dres, extra), accumulate=True)
        return self._identity_grad(dres, extra), self.c1.backward(d1)


RESNETS = {
    "resnet18": (BasicBlock, [2, 2, 2, 2]),
    "resnet34": (BasicBlock, [3, 4, 6, 3]),
    "resnet50": (Bottleneck, [3, 4, 6, 3]),
}
STEM_RP = 24    # im2col patch layout k = r*24 + s*3 + c (K.im2col_stem; the ResNet stem no longer uses it)
STEM_KP = 7 * STEM_RP


def _stem_maps():
    """index maps between torch [64,3,7,7] stem weights and the window layout [4, 64, 64] of the space-to-depth stem
    (K.stem_s2d + C.fprop_windows): tap t = s2d row offset, k = j*16 + (py*2 + px)*3 + c with j = s2d pixel inside the
    four-pixel window; filter row r = 2t + py - 1, filter column s = 2j + px - 1 (entries with r or s = -1, and the
    padding channels 12..15, carry no weight).  Returns (flat index into [3,7,7] per (t,k) or -1, validity mask)."""
    idx = np.full((4, 64), -1, np.int64)
    for t in range(4):
        for j in range(4):
            for py in range(2):
                for px in range(2):
                    r, s_ = 2 * t + py - 1, 2 * j + px - 1
                    if r < 0 or s_ < 0:
                        continue
                    for c in range(3):
                        idx[t, j * 16 + (py * 2 + px) * 3 + c] = (c * 7 + r) * 7 + s_
    return idx


_STEM_IDX = _stem_maps()


def stem_pack(w):
    """torch [64,3,7,7] -> [4, 64, 64] window layout (zeros where no filter tap lands)."""
    co = w.shape[0]
    idx = torch.from_numpy(_STEM_IDX).to(w.device)
    flat = torch.cat([w.reshape(co, 147), torch.zeros(co, 1, dtype=w.dtype, device=w.device)], 1)     # [-1] -> 0
    out = flat[:, idx.reshape(-1)].reshape(co, 4, 64)          # index -1 picks the appended zero
    return out.permute(1, 0, 2).contiguous()


def stem_unpack(d):
    """[4, 64, 64] (weights or their gradient) -> torch [64,3,7,7]; every filter tap appears exactly once"""
    co = d.shape[1]
    idx = torch.from_numpy(_STEM_IDX).to(d.device).reshape(-1)
    valid = idx >= 0
    out = torch.zeros(co, 147, dtype=d.dtype, device=d.device)
    out[:, idx[valid]] = d.permute(1, 0, 2).reshape(co, 256)[:, valid]
    return out.reshape(co, 3, 7, 7)


def stem_mask(device):
    """1 where a window-layout entry carries a filter tap, 0 elsewhere ([4, 1, 64]): the weight gradient of the window
    convolution is dense, the entries that are structurally zero must stay zero"""
    return torch.from_numpy((_STEM_IDX >= 0).astype(np.float32)).to(device).reshape(4, 1, 64)


class ResNetEncoder:
    """torchvision ResNet without the classifier; stage 5 dilated (stride 1, dilation 2) for output stride 16
    (smp `make_dilated(stage_list=[5], dilation_list=[2])`)."""

    def __init__(self, store, name, in_channels=3, dilated=True):
        assert in_channels == 3
        block, layers = RESNETS[name]

        def stem_init(shape):
            return stem_pack(kaiming_fan_out((64, 3, 7, 7)))
        self.stem_w = store.add("encoder.conv1.weight", (4, 64, 64), "conv_nt", stem_init)
        self.stem_mask = stem_mask(store.device)
        self.stem_bn = BatchNorm(store, "encoder.bn1", 64)
        self.blocks = []
        cin = 64
        self.out_channels = [3, 64]
        for li, (planes, n) in enumerate(zip([64, 128, 256, 512], layers)):
            stride = 1 if li == 0 else 2
            dil = 1
            if li == 3 and dilated:
                stride, dil = 1, 2
            stage = []
            for b in range(n):
                s = stride if b == 0 else 1
                ds = b == 0 and (li > 0 or cin != planes * block.expansion)
                stage.append(block(store, "encoder.layer%d.%d" % (li + 1, b), cin, planes, s, dil, ds))
                cin = planes * block.expansion
            self.blocks.append(stage)
            self.out_channels.append(cin)

    def stem_fprop(self, img, stats=None):
        """conv1 (7x7 / stride 2 / pad 3, 3 -> 64) as a 4x1 window convolution over the space-to-depth image: returns
        (window view kept for the weight gradient, pre-activation bf16 [N,H/2,W/2,64])"""
        n, _, h, w = img.shape
        _, xw = K.stem_s2d(img)
        pre = C.fprop_windows(xw, self.stem_w.bf16, 4, 1, h // 2, w // 2, stats=stats,
                              flops=2.0 * n * (h // 2) * (w // 2) * 64 * 147)       # the layer's own FLOPs
        return xw, pre

    def stem_wgrad(self, xw, dpre, out=None):
        """weight gradient of conv1 in the window layout [4,64,64] (structural zeros masked), accumulated into `out`"""
        n, ho, wo, _ = dpre.shape
        dw = C.wgrad_windows(xw, dpre, 4, 1, flops=2.0 * n * ho * wo * 64 * 147)
        dw.mul_(self.stem_mask)
        if out is None:
            return dw
        out.add_(dw)
        return out

    def forward(self, img, training):
        fused = training and FUSE_BN_STATS
        col, pre = self.stem_fprop(img, stats=self.stem_bn.stats_buffers() if fused else None)
        f1 = torch.empty_like(pre)
        self.stem_bn.forward(pre, f1, training, have_stats=fused)
        pooled, arg = K.maxpool_fwd(f1)
        feats = [f1]
        x = pooled
        for stage in self.blocks:
            for blk in stage:
                x = blk.forward(x, training)
            feats.append(x)
        self.ctx = (col, pre, f1, arg) if training else None
        return feats            # strides 2, 4, 8, 16, 16

    def backward(self, d_last, d_stride4=None, d_skips=None, on_ready=None):
        """d_last: gradient of the last feature map; d_stride4: gradient of the stride-4 map (DeepLab decoder
        skip); d_skips: gradients of feats[0..3] as returned by forward (UNet skips, None entries allowed).
        on_ready(offset): called when every parameter gradient at flat offset >= `offset` is final (a stage is done)."""
        col, pre, f1, arg = self.ctx
        self.ctx = None
        skips = list(d_skips) if d_skips is not None else [None, d_stride4, None, None]
        d = d_last
        for li in (3, 2, 1, 0):
            blocks = self.blocks[li]
            for bi in range(len(blocks) - 1, -1, -1):
                # the gradient of feats[li] (a decoder skip) joins the gradient of stage li's first block's input
                extra = skips[li] if (bi == 0 and li > 0) else None
                d = blocks[bi].backward(d, extra=extra)
            if on_ready is not None and li > 0:
                on_ready(first_offset(blocks))
        d = K.maxpool_bwd(fold_pair(d), arg, f1.shape)
        if skips[0] is not None:
            K.add_(d, skips[0])
        dpre = torch.empty_like(pre)
        self.stem_bn.backward(d, pre, f1, dpre)
        on_side(lambda: self.stem_wgrad(col, dpre, out=self.stem_w.grad), col, dpre)


class DepthwiseBN:
    """Conv2d(c, c, 3, stride, padding=dil, dilation=dil, groups=c, bias=False) -> BatchNorm2d -> ReLU6 (torchvision
    ConvBNActivation inside MobileNetV2's InvertedResidual)"""

    def __init__(self, store, conv_name, bn_name, c, stride, dil):
        self.dw = Depthwise3x3(store, conv_name, c, dil, stride, init=kaiming_fan_out)
        self.bn = BatchNorm(store, bn_name, c)

    def forward(self, x, training):
        pre = self.dw.forward(x, training)
        y = torch.empty_like(pre)
        self.bn.forward(pre, y, training, relu=True, relu6=True)
        self.ctx = pre if training else None
        return y

    def backward(self, dy):
        pre = self.ctx
        self.ctx = None
        dpre = torch.empty_like(pre)
        self.bn.backward(dy, pre, None, dpre, relu=True, relu6=True)
        return self.dw.backward(dpre)


class InvertedResidual:
    """torchvision MobileNetV2 block: [1x1 expand + BN + ReLU6] -> depthwise 3x3 + BN + ReLU6 -> 1x1 project + BN (linear),
    identity added when stride == 1 and cin == cout.  state_dict names: features.i.conv.{0.0,0.1,1.0,1.1,2,3}
    (expand ratio 1: conv.{0.0,0.1,1,2})."""

    def __init__(self, store, name, cin, cout, stride, expand, dil=1):
        hid = cin * expand
        n = name + ".conv."
        self.expand = ConvBN(store, n + "0.0", n + "0.1", cin, hid, 1, relu="relu6") if expand != 1 else None
        i = 1 if expand != 1 else 0
        self.dw = DepthwiseBN(store, n + "%d.0" % i, n + "%d.1" % i, hid, stride, dil)
        self.project = ConvBN(store, n + "%d" % (i + 1), n + "%d" % (i + 2), hid, cout, 1, relu=False)
        self.use_res = stride == 1 and cin == cout

    def forward(self, x, training):
        h = self.expand.forward(x, training) if self.expand else x
        return self.project.forward(self.dw.forward(h, training), training, res=x if self.use_res else None)

    def backward(self, dy):
        """the projection is linear, so the identity branch's gradient is dy itself: the conv branch accumulates into it
        (dy is this block's private copy of the gradient: nothing else reads it afterwards)"""
        d = self.dw.backward(self.project.backward(dy))
        if self.expand is None:
            if self.use_res:
                K.add_(d, dy)
            return d
        if self.use_res:
            return self.expand.backward(d, dx=dy, accumulate=True)
        return self.expand.backward(d)


MBV2_SETTING = [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]
MBV2_STEM_RP = 16     # 3x3 stem patches: k = r*16 + s*3 + c (9 values + 7 zeros per filter row)


def stem_pack_rows(w, rp):
    """torch [Cout,3,R,S] -> [1,Cout,R*rp] in the row-pitched patch order of K.im2col_stem(row_pitch=rp)."""
    co, _, r, s_ = w.shape
    v = w.permute(0, 2, 3, 1).reshape(co, r, 3 * s_)
    return torch.cat([v, torch.zeros(co, r, rp - 3 * s_, dtype=w.dtype, device=w.device)], 2).reshape(1, co, r * rp)


def stem_unpack_rows(d, r, s_, rp):
    co = d.shape[1]
    return d[0].reshape(co, r, rp)[:, :, :3 * s_].reshape(co, r, s_, 3).permute(0, 3, 1, 2).contiguous()


class MobileNetV2Encoder:
    """torchvision mobilenet_v2 `features` (the reference's only reachable backbone, models/__init__.py:16) as smp wraps
    it: stages features[:2], [2:4], [4:7], [7:14], [14:]; out_channels (3, 16, 24, 32, 96, 1280); output stride 16 =
    every conv of the last stage at stride 1 / dilation 2 (smp make_dilated)."""

    def __init__(self, store, in_channels=3, dilated=True):
        assert in_channels == 3
        f = "encoder.features."
        self.stem_w = store.add(f + "0.0.weight", (1, 32, 3 * MBV2_STEM_RP), "conv_nt",
                                lambda s: stem_pack_rows(kaiming_fan_out((32, 3, 3, 3)), MBV2_STEM_RP))
        self.stem_bn = BatchNorm(store, f + "0.1", 32)
        self.blocks = []          # (feature index, block)
        cin, idx = 32, 1
        for t, c, n, s_ in MBV2_SETTING:
            for i in range(n):
                stride = s_ if i == 0 else 1
                dil = 1
                if dilated and idx >= 14:
                    stride, dil = 1, 2
                self.blocks.append((idx, InvertedResidual(store, f + str(idx), cin, c, stride, t, dil)))
                cin = c
                idx += 1
        self.last = ConvBN(store, f + "18.0", f + "18.1", cin, 1280, 1, relu="relu6")
        self.out_channels = [3, 16, 24, 32, 96, 1280]
        self.stage_ends = (1, 3, 6, 13)       # feature index after which a stage output is taken (then the last conv)

    def forward(self, img, training):
        col = K.im2col_stem(img, 3, 3, 2, 1, 3 * MBV2_STEM_RP, row_pitch=MBV2_STEM_RP)
        fused = training and FUSE_BN_STATS
        pre = C.fprop(col, self.stem_w.bf16, 1, 1, stats=self.stem_bn.stats_buffers() if fused else None)
        x = torch.empty_like(pre)
        self.stem_bn.forward(pre, x, training, relu=True, relu6=True, have_stats=fused)
        feats = []
        for idx, blk in self.blocks:
            x = blk.forward(x, training)
            if idx in self.stage_ends:
                feats.append(x)
        feats.append(self.last.forward(x, training))
        self.ctx = (col, pre) if training else None
        return feats            # strides 2, 4, 8, 16, 16

    def backward(self, d_last, d_stride4=None, d_skips=None, on_ready=None):
        col, pre = self.ctx
        self.ctx = None
        skips = list(d_skips) if d_skips is not None else [None, d_stride4, None, None]
        d = self.last.backward(d_last)
        for idx, blk in reversed(self.blocks):
            if idx in self.stage_ends:
                sk = skips[self.stage_ends.index(idx)]
                if sk is not None:
                    K.add_(d, sk)
                if on_ready is not None and idx == 13:      # features.14 .. 18 are done (most of the encoder's weights)
                    on_ready(first_offset([b for i, b in self.blocks if i > idx]))
            d = blk.backward(d)
        dpre = torch.empty_like(pre)
        self.stem_bn.backward(d, pre, None, dpre, relu=True, relu6=True)
        on_side(lambda: C.wgrad(col, dpre, 1, 1, 1, 0, 1, out=self.stem_w.grad), col, dpre)


class SeparableConvBN:
    """smp SeparableConv2d(depthwise 3x3 + pointwise 1x1, bias=False) -> BN -> ReLU"""

    def __init__(self, store, dw_name, pw_name, bn_name, cin, cout, dil):
        self.dw = Depthwise3x3(store, dw_name, cin, dil)
        self.pw = ConvBN(store, pw_name, bn_name, cin, cout, 1, init=kaiming_uniform_default)

    def forward(self, x, training, out=None):
        return self.pw.forward(self.dw.forward(x, training), training, out=out)

    def backward(self, dy, dx=None, accumulate=False):
        return self.dw.backward(self.pw.backward(dy), dx=dx, accumulate=accumulate)


class DeepLabV3PlusDecoder:
    def __init__(self, store, enc_channels, out_channels=256, rates=(12, 24, 36)):
        cenc, chigh = enc_channels[-1], enc_channels[-4]
        oc = out_channels
        u = kaiming_uniform_default
        a = "decoder.aspp.0."
        self.b0 = ConvBN(store, a + "convs.0.0", a + "convs.0.1", cenc, oc, 1, init=u)
        self.br = [SeparableConvBN(store, a + "convs.%d.0.0" % (i + 1), a + "convs.%d.0.1" % (i + 1),
                                   a + "convs.%d.1" % (i + 1), cenc, oc, r) for i, r in enumerate(rates)]
        self.bp = ConvBN(store, a + "convs.4.1", a + "convs.4.2", cenc, oc, 1, init=u)
        self.project = ConvBN(store, a + "project.0", a + "project.1", 5 * oc, oc, 1, init=u)
        self.sep = SeparableConvBN(store, "decoder.aspp.1.0", "decoder.aspp.1.1", "decoder.aspp.2", oc, oc, 1)
        self.block1 = ConvBN(store, "decoder.block1.0", "decoder.block1.1", chigh, 48, 1, init=u)
        self.block2 = SeparableConvBN(store, "decoder.block2.0.0", "decoder.block2.0.1", "decoder.block2.1", 48 + oc,
                                      oc, 1)
        self.oc = oc

    def forward(self, feats, training, dropout_seed):
        x, high = feats[-1], feats[-4]
        n, h, w, cenc = x.shape
        oc = self.oc
        cat = torch.empty((n, h, w, 5 * oc), dtype=BF16, device=x.device)
        self.b0.forward(x, training, out=cat[..., 0:oc])
        for i, br in enumerate(self.br):
            br.forward(x, training, out=cat[..., (i + 1) * oc:(i + 2) * oc])
        pooled = K.f32_to_bf16(K.global_sum(x, 1.0 / (h * w))).view(n, 1, 1, cenc)
        pv = self.bp.forward(pooled, training)
        K.broadcast_pixels(pv, cat[..., 4 * oc:5 * oc])
        pj = self.project.forward(cat, training, dropout_seed=dropout_seed if training else None)
        a = self.sep.forward(pj, training)
        hh, hw = high.shape[1:3]
        cat2 = torch.empty((n, hh, hw, oc + 48), dtype=BF16, device=x.device)
        K.upsample_fwd(a, cat2[..., 0:oc])
        self.block1.forward(high, training, out=cat2[..., oc:oc + 48])
        out = self.block2.forward(cat2, training)
        self.ctx = (x.shape, a.shape, cat.shape, cat2.shape) if training else None
        return out

    def backward(self, dy):
        xs, ash, cats, cat2s = self.ctx
        self.ctx = None
        n, h, w, cenc = xs
        oc = self.oc
        dev = dy.device
        dcat2 = self.block2.backward(dy)                                  # [n, hh, hw, oc+48]
        d_high = self.block1.backward(dcat2[..., oc:oc + 48])
        da = torch.empty(ash, dtype=BF16, device=dev)
        K.upsample_bwd(dcat2[..., 0:oc], da)
        dpj = self.sep.backward(da)
        dcat = self.project.backward(dpj)                                 # [n, h, w, 5*oc]
        dx = self.b0.backward(dcat[..., 0:oc])
        for i, br in enumerate(self.br):       # each branch's depthwise data gradient lands on dx directly
            br.backward(dcat[..., (i + 1) * oc:(i + 2) * oc], dx=dx, accumulate=True)
        dpv = K.f32_to_bf16(K.global_sum(dcat[..., 4 * oc:5 * oc], 1.0)).view(n, 1, 1, oc)
        dpooled = self.bp.backward(dpv)                                   # [n,1,1,cenc]
        K.broadcast_add_pixels(dpooled.reshape(n, cenc).float().contiguous(), dx, 1.0 / (h * w))
        return dx, d_high


class UnetDecoder:
    """smp 0.2.0 UnetDecoder(decoder_channels=(256,128,64,32,16), use_batchnorm=True, center=False): five
    blocks of [nearest x2 up-sampling, concat skip, Conv3x3-BN-ReLU, Conv3x3-BN-ReLU]."""

    def __init__(self, store, enc_channels, decoder_channels=(256, 128, 64, 32, 16)):
        enc = list(enc_channels[1:])[::-1]                   # [f5, f4, f3, f2, f1] channels
        in_ch = [enc[0]] + list(decoder_channels[:-1])
        skip_ch = enc[1:] + [0]
        self.blocks = []
        for i, (ic, sc, oc) in enumerate(zip(in_ch, skip_ch, decoder_channels)):
            n = "decoder.blocks.%d." % i
            c1 = ConvBN(store, n + "conv1.0", n + "conv1.1", ic + sc, oc, 3, 1, 1, 1, init=kaiming_uniform_default)
            c2 = ConvBN(store, n + "conv2.0", n + "conv2.1", oc, oc, 3, 1, 1, 1, init=kaiming_uniform_default)
            self.blocks.append((c1, c2, ic, sc))
        self.out_channels = decoder_channels[-1]

    def forward(self, feats, training):
        """feats = [f1, f2, f3, f4, f5] (strides 2..32)"""
        x = feats[-1]
        skips = feats[:-1][::-1] + [None]
        for (c1, c2, ic, sc), skip in zip(self.blocks, skips):
            n, h, w, _ = x.shape
            cat = torch.empty((n, 2 * h, 2 * w, ic + sc), dtype=BF16, device=x.device)
            K.nearest2x_fwd(x, cat[..., :ic])
            if skip is not None:
                K.copy_(skip, cat[..., ic:])
            x = c2.forward(c1.forward(cat, training), training)
        return x

    def backward(self, dy):
        """returns (gradient of f5, [gradients of f1..f4])"""
        d = dy
        d_skips = []
        for (c1, c2, ic, sc) in reversed(self.blocks):
            dcat = c1.backward(c2.backward(d))
            n, h2, w2, _ = dcat.shape
            if sc:
                d_skips.append(dcat[..., ic:])
            d = torch.empty((n, h2 // 2, w2 // 2, ic), dtype=BF16, device=dcat.device)
            K.nearest2x_bwd(dcat[..., :ic], d)
        # d_skips was collected for f1, f2, f3, f4 in that order (blocks reversed: last block has no skip)
        return d, d_skips


class _SegNetFunction(torch.autograd.Function):
    """`logits, pooled = model(x)` with an autograd node behind it, so that the reference's own lines
    `seg_loss.backward(); model_optimizer.step()` (search_dg.py:170-172) drive the engine: backward receives
    d(loss)/d(logits) (and, if the pooled feature was used undetached, its gradient), runs the engine's explicit backward
    pass and leaves the parameter gradients in the flat buffer that the `.grad` of every `model.parameters()` leaf views."""

    @staticmethod
    def forward(ctx, net, x, *leaves):
        dec, pooled = net.features(x)
        z = net._head(dec)
        n, _, hh, ww = x.shape
        logits = torch.empty((n, net.classes, hh, ww), dtype=torch.float32, device=x.device)
        K.seg_loss_fwd(z, torch.zeros_like(logits), 0.5, torch.zeros(1, dtype=torch.float64, device=x.device),
                       torch.zeros((n, net.classes, 3), dtype=torch.int32, device=x.device), logits)
        ctx.net, ctx.dec, ctx.z_shape = net, dec, tuple(z.shape)
        ctx.set_materialize_grads(False)
        return logits, pooled

    @staticmethod
    def backward(ctx, dlogits, dpooled):
        net, dec = ctx.net, ctx.dec
        net._bind_grads()
        if dlogits is None:
            dlogits = torch.zeros((dec.shape[0], net.classes, ctx.z_shape[1] * 4, ctx.z_shape[2] * 4), device=dec.device)
        dz = K.upsample_logits_bwd(dlogits.float(), ctx.z_shape)
        _Side.armed = True
        ddec = torch.empty(dec.shape, dtype=BF16, device=dec.device)

        def add_pooled(d_last):
            if dpooled is not None:       # pooled = mean over pixels of the last encoder map
                hw = d_last.shape[1] * d_last.shape[2]
                K.broadcast_add_pixels(dpooled.float().contiguous(), d_last, 1.0 / hw)
            return d_last
        if net.arch == "unet":
            K.seg_head3x3_bwd(dz, dec, net.head_w.data, ddec, net.head_w.grad, net.head_b.grad)
            d_last, d_skips = net.decoder.backward(ddec)
            net.encoder.backward(add_pooled(d_last), d_skips=d_skips)
        else:
            K.seg_head_bwd(dz, dec, net.head_w.data, ddec, net.head_w.grad, net.head_b.grad)
            d_last, d_high = net.decoder.backward(ddec)
            net.encoder.backward(add_pooled(d_last), d_high)
        join_side()
        _Side.armed = False
        net.steps += 1
        if net.dropout_enabled:
            net.seed_dev += 1
        return (None, None) + (None,) * len(net._leaves)


class SegNet:
    """encoder + decoder + segmentation head (Conv2d(256, classes, 1) -> UpsamplingBilinear2d(4)) and the
    patched classification head (AdaptiveAvgPool2d(1) + flatten of the last encoder map, models/heads.py:14-25)."""

    def __init__(self, encoder_name="resnet50", classes=2, device="cuda", seed=0, arch="deeplabv3plus"):
        if not torch.cuda.is_available():
            raise RuntimeError("aadg_b200.nn needs a CUDA device: there is no CPU path")
        self.device = torch.device(device)
        gen_state = torch.random.get_rng_state()
        torch.manual_seed(seed)
        self.store = ParamStore(self.device)
        self.arch = arch
        self.classes = classes
        def make_encoder(dilated):
            if encoder_name == "mobilenet_v2":
                return MobileNetV2Encoder(self.store, dilated=dilated)
            return ResNetEncoder(self.store, encoder_name, dilated=dilated)
        if arch == "unet":
            self.encoder = make_encoder(False)
            self.decoder = UnetDecoder(self.store, self.encoder.out_channels)
            hc = self.decoder.out_channels
            self.head_w = self.store.add(
                "segmentation_head.0.weight", (classes, 9, hc), "f32",
                lambda s: kaiming_uniform_default((classes, hc, 3, 3)).permute(0, 2, 3, 1).reshape(classes, 9, hc))
            bound = 1.0 / math.sqrt(hc * 9)
        else:
            self.encoder = make_encoder(True)
            self.decoder = DeepLabV3PlusDecoder(self.store, self.encoder.out_channels)
            self.head_w = self.store.add("segmentation_head.0.weight", (classes, 256), "f32",
                                         lambda s: kaiming_uniform_default((classes, 256, 1, 1)).reshape(classes, 256))
            bound = 1.0 / math.sqrt(256)
        self.head_b = self.store.add("segmentation_head.0.bias", (classes,), "f32",
                                     lambda s: (torch.rand(s) * 2 - 1) * bound)
        self.store.finalize()
        torch.random.set_rng_state(gen_state)
        # torch.nn.Module-like parameter surface: one leaf tensor per parameter, a VIEW of the flat fp32 master buffer,
        # whose .grad is a view of the flat gradient buffer -- torch.optim.Adam(model.parameters()) updates the masters
        # in place (the bf16 weight copies are refreshed lazily: ParamStore.sync())
        self._leaves = [p.data.requires_grad_(True) for p in self.store.items]
        self.training = True
        self.dropout_seed = 0x5EED0000 + seed
        self.dropout_enabled = True      # Dropout(0.5) of the ASPP projection (train mode)
        self.steps = 0
        # the current dropout seed (dropout_seed + steps) in device memory: the kernels read it there (flag 64), so a
        # captured CUDA graph draws a fresh mask on every replay
        self.seed_dev = torch.full((1,), self.dropout_seed, dtype=torch.int64, device=self.device)

    # ---- torch.nn.Module-like surface ---------------------------------------------------------------
    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def cuda(self, *a, **k):
        return self

    def parameters(self):
        return list(self._leaves)

    def named_parameters(self):
        return [(p.name, leaf) for p, leaf in zip(self.store.items, self._leaves)]

    def zero_grad(self, set_to_none=True):
        for leaf in self._leaves:
            leaf.grad = None
        self.store.zero_grad()

    def _bind_grads(self):
        """before a backward pass through the autograd surface: every leaf's .grad is its slice of the flat gradient
        buffer; after `optimizer.zero_grad()` (set_to_none, the default) the buffer starts from zero, otherwise the
        new gradients accumulate onto the old ones like torch's"""
        if all(leaf.grad is None for leaf in self._leaves):
            self.store.grads.zero_()
        for p, leaf in zip(self.store.items, self._leaves):
            if leaf.grad is None:
                leaf.grad = p.grad
            elif leaf.grad.data_ptr() != p.grad.data_ptr():
                raise RuntimeError("parameter %s: .grad was replaced by a foreign tensor" % p.name)

    def named_params(self):
        return {p.name: p for p in self.store.items}

    def _bns(self):
        out = []

        def walk(o):
            if isinstance(o, BatchNorm):
                out.append(o)
            elif isinstance(o, (list, tuple)):
                for i in o:
                    walk(i)
            elif hasattr(o, "__dict__") and not isinstance(o, (ParamStore, Param, torch.Tensor)):
                for v in vars(o).values():
                    walk(v)
        walk(self.encoder)
        walk(self.decoder)
        return out

    def state_dict(self):
        """smp 0.2.0 key names and torch weight layouts."""
        sd = {}
        for p in self.store.items:
            d = p.data.detach().clone()
            if p.name == "encoder.conv1.weight":
                d = stem_unpack(d)
            elif p.name == "encoder.features.0.0.weight":
                d = stem_unpack_rows(d, 3, 3, MBV2_STEM_RP)
            elif p.kind in ("conv", "conv_nt"):
                k = int(round(math.sqrt(p.shape[0])))
                d = from_taps(d, k, k)
            elif p.name.startswith("segmentation_head.0.weight"):
                d = d.reshape(self.classes, 256, 1, 1) if self.arch != "unet" else \
                    d.reshape(self.classes, 3, 3, -1).permute(0, 3, 1, 2).contiguous()
            elif len(p.shape) == 2 and p.shape[0] == 9:
                d = d.t().reshape(p.shape[1], 1, 3, 3).contiguous()
            sd[p.name] = d
        for bn in self._bns():
            sd[bn.name + ".running_mean"] = bn.running_mean.clone()
            sd[bn.name + ".running_var"] = bn.running_var.clone()
            sd[bn.name + ".num_batches_tracked"] = torch.tensor(bn.num_batches_tracked)
        return sd

    def load_state_dict(self, sd, strict=True):
        named = self.named_params()
        missing = [k for k in named if k not in sd]
        if strict and missing:
            raise KeyError("missing keys: %s" % missing[:5])
        for name, p in named.items():
            if name not in sd:
                continue
            w = sd[name].detach().to(self.device, torch.float32)
            if name == "encoder.conv1.weight":
                w = stem_pack(w)
            elif name == "encoder.features.0.0.weight":
                w = stem_pack_rows(w, MBV2_STEM_RP)
            elif p.kind in ("conv", "conv_nt"):
                w = to_taps(w)
            elif name == "segmentation_head.0.weight":
                w = w.reshape(self.classes, 256) if self.arch != "unet" else \
                    w.permute(0, 2, 3, 1).reshape(self.classes, 9, -1)
            elif len(p.shape) == 2 and p.shape[0] == 9:
                w = w.reshape(p.shape[1], 9).t().contiguous()
            p.data.copy_(w.reshape(p.shape))
        for bn in self._bns():
            if bn.name + ".running_mean" in sd:
                bn.running_mean.copy_(sd[bn.name + ".running_mean"])
                bn.running_var.copy_(sd[bn.name + ".running_var"])
        self.store.refresh()

    # ---- forward / backward ---------------------------------------------------------------------------
    def features(self, x):
        """encoder + decoder; returns (decoder map bf16 [N,H/4,W/4,256], pooled encoder feature fp32 [N,C])."""
        if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 3):
            raise ValueError("input must be CUDA float32 [N,3,H,W]")
        self.store.sync()
        feats = self.encoder.forward(x.contiguous(), self.training)
        last = feats[-1]
        pooled = K.global_sum(last, 1.0 / (last.shape[1] * last.shape[2]))
        if self.arch == "unet":
            dec = self.decoder.forward(feats, self.training)
        else:
            seed = self.seed_dev if (self.training and self.dropout_enabled) else None
            dec = self.decoder.forward(feats, self.training, seed)
        return dec, pooled

    def _head(self, dec):
        if self.arch == "unet":
            return K.seg_head3x3_fwd(dec, self.head_w.data, self.head_b.data)
        return K.seg_head_fwd(dec, self.head_w.data, self.head_b.data)

    def __call__(self, x):
        """smp call shape: (logits float32 [N,classes,H,W], pooled feature float32 [N,C_enc]).  In training mode with
        autograd enabled the outputs carry a graph (see _SegNetFunction): `loss.backward()` runs the engine's backward."""
        if self.training and torch.is_grad_enabled():
            return _SegNetFunction.apply(self, x.detach(), *self._leaves)
        dec, pooled = self.features(x)
        z = self._head(dec)
        n, _, hh, ww = x.shape
        logits = torch.empty((n, self.classes, hh, ww), dtype=torch.float32, device=x.device)
        dummy_t = torch.zeros_like(logits)
        K.seg_loss_fwd(z, dummy_t, 0.5, torch.zeros(1, dtype=torch.float64, device=x.device),
                       torch.zeros((n, self.classes, 3), dtype=torch.int32, device=x.device), logits)
        self.ctx = (dec, z) if self.training else None
        return logits, pooled

    forward = __call__

    def evaluate_batch(self, x, target, thr=0.75):
        """Inference + metrics in one pass (validate(), search_dg.py:236-245): logits float32 [N,classes,H,W], the BCE
        of sigmoid(logits) vs target (mean) and the TP/FP/FN counts of (sigmoid(logits) > thr) per sample and class."""
        dec, pooled = self.features(x)
        z = self._head(dec)
        n, _, hh, ww = x.shape
        logits = torch.empty((n, self.classes, hh, ww), dtype=torch.float32, device=x.device)
        loss_sum = torch.zeros(1, dtype=torch.float64, device=x.device)
        counts = torch.zeros((n, self.classes, 3), dtype=torch.int32, device=x.device)
        K.seg_loss_fwd(z, target, thr, loss_sum, counts, logits)
        return dict(logits=logits, counts=counts, pooled=pooled, loss=(loss_sum / float(logits.numel())).float())

    def loss_step(self, x, target, thr=0.5, want_logits=False, on_ready=None):
        """Forward + BCELoss(sigmoid(logits), target) (mean) + Dice counts, then the whole backward.
        Gradients are left in store.grads (call store.zero_grad() before, store.adam_step() after).
        on_ready(offset): called during the backward pass each time every gradient at flat offset >= `offset` is final
        (head + decoder, then encoder stage by stage, finally 0) -- the hook a bucketed gradient all-reduce overlaps on.
        Returns dict(loss [1] fp32 tensor, counts int32 [N,classes,3], pooled fp32 [N,C], logits or None)."""
        assert self.training
        n, _, hh, ww = x.shape
        if on_ready is not None:        # a gradient is final once the side stream's weight gradients have joined
            user_ready = on_ready

            def on_ready(offset):
                join_side()
                user_ready(offset)
        dec, pooled = self.features(x)
        z = self._head(dec)
        loss_sum = torch.zeros(1, dtype=torch.float64, device=x.device)
        counts = torch.zeros((n, self.classes, 3), dtype=torch.int32, device=x.device)
        logits = torch.empty((n, self.classes, hh, ww), dtype=torch.float32, device=x.device) if want_logits else None
        K.seg_loss_fwd(z, target, thr, loss_sum, counts, logits)
        numel = float(n * self.classes * hh * ww)
        dz = K.seg_loss_bwd(z, target, 1.0 / numel)
        _Side.armed = True
        ddec = torch.empty(dec.shape, dtype=BF16, device=x.device)
        if self.arch == "unet":
            K.seg_head3x3_bwd(dz, dec, self.head_w.data, ddec, self.head_w.grad, self.head_b.grad)
            d_last, d_skips = self.decoder.backward(ddec)
            if on_ready is not None:
                on_ready(first_offset(self.decoder))
            self.encoder.backward(d_last, d_skips=d_skips, on_ready=on_ready)
        else:
            K.seg_head_bwd(dz, dec, self.head_w.data, ddec, self.head_w.grad, self.head_b.grad)
            d_last, d_high = self.decoder.backward(ddec)
            if on_ready is not None:
                on_ready(first_offset(self.decoder))
            self.encoder.backward(d_last, d_high, on_ready=on_ready)
        join_side()
        _Side.armed = False
        if on_ready is not None:
            on_ready(0)
        self.steps += 1
        if self.dropout_enabled:
            self.seed_dev += 1
        return dict(loss=(loss_sum / numel).float(), counts=counts, pooled=pooled, logits=logits)


class GraphedTrainStep:
    """zero_grad -> loss_step (forward, loss, backward, with the gradient-ready hooks) -> `before_update()` -> Adam from
    device-resident scalars, captured ONCE into a CUDA graph and replayed: ~2 000 kernel launches become one
    cudaGraphLaunch.  `images` / `labels` are the static input buffers the caller refills before every replay(); the
    returned dict holds static output tensors (overwritten by the next replay).  Everything that changes from step to
    step lives in device memory (dropout seed, Adam step count and rate), nothing is baked into the graph by value."""

    def __init__(self, model, images, labels, on_ready=None, before_update=None, pool=None):
        assert C.TIMING is None, "per-launch event timing cannot be captured"
        self.model = model
        store = model.store
        bns = model._bns()
        saved = (model.steps, store.step, [bn.num_batches_tracked for bn in bns])
        torch.cuda.synchronize(model.device)
        self.graph = torch.cuda.CUDAGraph()
        from .. import _lib
        calls0 = _lib.CALLS
        with torch.cuda.graph(self.graph, pool=pool):
            store.zero_grad()
            self.out = model.loss_step(images, labels, on_ready=on_ready)
            if before_update is not None:
                before_update()
            store.adam_step_dev()
        self.launches = _lib.CALLS - calls0        # C-ABI launches inside the graph (replayed every step)
        # capture executes nothing on the device: take back the host-side counters it advanced
        model.steps, store.step = saved[0], saved[1]
        for bn, nb in zip(bns, saved[2]):
            bn.num_batches_tracked = nb
        self._bns = bns

    def replay(self):
        from .. import _lib
        self.graph.replay()
        _lib.CALLS += self.launches
        self.model.steps += 1
        self.model.store.step += 1
        for bn in self._bns:
            bn.num_batches_tracked += 1
        return self.out


def dice_from_counts(counts):
    """torchmetrics F1(num_classes=2, average=None, mdmc_average='samplewise')[1] per class:
    mean over samples of 2TP/(2TP+FP+FN), 0 on 0/0.  counts int [N,K,3] -> [K]."""
    c = counts.to(torch.float64)
    den = 2 * c[..., 0] + c[..., 1] + c[..., 2]
    f1 = torch.where(den > 0, 2 * c[..., 0] / den.clamp(min=1), torch.zeros_like(den))
    return f1.mean(0)


def DeepLabV3Plus(encoder_name="resnet50", encoder_depth=5, encoder_weights=None, encoder_output_stride=16,
                  decoder_channels=256, decoder_atrous_rates=(12, 24, 36), in_channels=3, classes=1, activation=None,
                  upsampling=4, aux_params=None, device="cuda", seed=0):
    """smp.DeepLabV3Plus constructor shape (models/__init__.py:17-23).  `encoder_weights` must be None
    (no network for ImageNet checkpoints: load_state_dict() accepts smp-named weights instead)."""
    if encoder_name not in RESNETS and encoder_name != "mobilenet_v2":
        raise NotImplementedError("encoder %r: mobilenet_v2 and resnet18/34/50 are implemented" % encoder_name)
    if encoder_weights not in (None, "none"):
        raise NotImplementedError("pretrained encoder weights cannot be downloaded here; use load_state_dict()")
    if (encoder_depth, encoder_output_stride, decoder_channels, tuple(decoder_atrous_rates), in_channels, upsampling,
            activation) != (5, 16, 256, (12, 24, 36), 3, 4, None):
        raise NotImplementedError("only the reference's DeepLabV3+ configuration is implemented")
    if aux_params is not None and aux_params.get("pooling", "avg") != "avg":
        raise NotImplementedError("aux head: average pooling only (models/heads.py)")
    return SegNet(encoder_name, classes, device, seed)


def Unet(encoder_name="resnet34", encoder_depth=5, encoder_weights=None, decoder_use_batchnorm=True,
         decoder_channels=(256, 128, 64, 32, 16), decoder_attention_type=None, in_channels=3, classes=1,
         activation=None, aux_params=None, device="cuda", seed=0):
    """smp.Unet constructor shape (the north star's UNet / RVS configuration)."""
    if encoder_name not in RESNETS and encoder_name != "mobilenet_v2":
        raise NotImplementedError("encoder %r: mobilenet_v2 and resnet18/34/50 are implemented" % encoder_name)
    if encoder_weights not in (None, "none"):
        raise NotImplementedError("pretrained encoder weights cannot be downloaded here; use load_state_dict()")
    if (encoder_depth, decoder_use_batchnorm, tuple(decoder_channels), decoder_attention_type, in_channels,
            activation) != (5, True, (256, 128, 64, 32, 16), None, 3, None):
        raise NotImplementedError("only smp's default Unet configuration is implemented")
    return SegNet(encoder_name, classes, device, seed, arch="unet")
