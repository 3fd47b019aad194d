"""One step of the reference's search loop (`search_dg.train`, search_dg.py:123-206) on the CUDA engine.

    augment (DGMultiPolicy + DGRandomScaleCrop + Normalize_dg/ToTensor, uint8 bank)  ->  model forward  ->  momentum
    discriminator features  ->  Sinkhorn diversity rewards  ->  BCE backward + Adam  ->  discriminator step

Row order of every per-image tensor is the reference's collate order (b*D + d)*M + j
(data/transform.py:323-340).

Multi-GPU (torch.distributed initialised; one process per GPU): each rank owns a contiguous shard of the source images
(`shard_sources`).  Two exchanges, both NCCL:
  * the flat fp32 gradient buffer is all-reduced in BUCKETS, each issued (async, on NCCL's own stream) as soon as the
    backward pass has finished the stage that owns it -- head + decoder first, the encoder stage by stage -- so the
    transfers overlap the rest of the backward pass (the DDP bucketing of models/__init__.py:39); the 1/world
    averaging is folded into the Adam kernel's gradient load;
  * the 128-d momentum-discriminator features and the domain codes travel in ONE all_gather_into_tensor, after which
    every rank runs the same fused reward kernel => identical rewards on every rank.
Decisions (sub-policy choice, Cutout centres, scale/crop, soft domain labels) are counter-based and keyed by the GLOBAL
source index, so a sharded run makes the decisions a single-process run of the whole batch makes.

CUDA graph: the model part of the step (zero_grad, forward, loss, backward with the bucketed all-reduce, Adam) is
captured once per input shape and replayed (`graph=True`); everything that varies per step is read from device memory.
"""
import numpy as np
import torch
import torch.distributed as dist

from ..data import decisions as D
from ..nn import network as NW
from ..nn.network import dice_from_counts
from ..ops import sinkhorn as SK
from ..ops import u8 as U8
from .discriminator import MomentumFeatureDiscriminator
from .losses import CrossEntropy


def gather_rows(*tensors):
    """all-gather row blocks rank-major in ONE collective: the [n, c_i] float32 tensors are packed side by side, every
    rank ends with the same [world*n, c_i] tensors (the cross-domain feature exchange feeding the Sinkhorn reward)."""
    world = dist.get_world_size()
    packed = torch.cat([t.reshape(t.shape[0], -1).float() for t in tensors], dim=1).contiguous()
    out = torch.empty((world * packed.shape[0], packed.shape[1]), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, packed)
    res, c0 = [], 0
    for t in tensors:
        c = t.reshape(t.shape[0], -1).shape[1]
        res.append(out[:, c0:c0 + c].contiguous())
        c0 += c
    return res


def average_(flat):
    """gradient all-reduce (sum) and division by the world size, in place."""
    dist.all_reduce(flat)
    flat.mul_(1.0 / dist.get_world_size())
    return flat


def shutdown(engines=(), timeout_s=20.0):
    """orderly end of a distributed run: captured graphs released, device drained, process group destroyed.  NCCL's
    communicator teardown has been seen to wait forever when kernels of the communicator were captured into a CUDA graph,
    so the WHOLE teardown runs under a watchdog: if it has not finished after `timeout_s` the process exits hard with
    status 0 (every result has been printed by then; the OS reclaims the device)."""
    import gc
    import os
    import sys
    import threading
    if not (dist.is_available() and dist.is_initialized()):
        for e in engines:
            e.close()
        return

    def _bail():
        sys.stdout.flush()
        sys.stderr.write("aadg_b200: distributed teardown did not finish within %.0f s; exiting\n" % timeout_s)
        sys.stderr.flush()
        os._exit(0)
    watchdog = threading.Timer(timeout_s, _bail)
    watchdog.daemon = True
    watchdog.start()
    for e in engines:
        e.close()
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    dist.destroy_process_group()
    watchdog.cancel()


def shard_sources(n_sources, rank, world):
    """contiguous block of source-image indices owned by `rank` (units are independent)."""
    per = (n_sources + world - 1) // world
    return list(range(min(rank * per, n_sources), min((rank + 1) * per, n_sources)))


class BucketedAllReduce:
    """The gradient exchange of one step: `ready(offset)` (called by the backward pass when every gradient at flat
    offset >= offset is final) launches an async all-reduce of the new slice [offset, previous offset); `wait()` joins
    them all before the optimiser reads the buffer.  Slices smaller than `min_elems` are merged into the next one."""

    def __init__(self, flat, min_elems=1 << 20, group=None):
        self.flat, self.min_elems, self.group = flat, min_elems, group
        self.hi = flat.numel()
        self.works = []

    def ready(self, offset):
        offset = int(offset)
        if offset > 0 and self.hi - offset < self.min_elems:
            return
        if offset < self.hi:
            self.works.append(dist.all_reduce(self.flat[offset:self.hi], group=self.group, async_op=True))
            self.hi = offset

    def wait(self):
        self.ready(0)
        for w in self.works:
            w.wait()
        self.works = []
        self.hi = self.flat.numel()


class SearchEngine:
    def __init__(self, model, n_domains=3, M=6, lr=1e-3, weight_decay=0.0, dataset="optic", seed=1023,
                 crop=None, scale_range=(1, 1.5), graph=False, n_sources_total=None, src_offset=None,
                 distributed=None):
        """graph: capture the model part of the step into a CUDA graph (from the second step on).
        distributed: None = use torch.distributed when it is initialised; False = single-process engine regardless.
        n_sources_total / src_offset: this rank's sources are src_offset.. of a global batch of n_sources_total
        (default: equal shards in rank order); they key the decision draws."""
        self.model = model
        self.M, self.n_domains, self.dataset = M, n_domains, dataset
        self.lr, self.wd = lr, weight_decay
        dev = model.device
        enc_c = model.encoder.out_channels[-1]
        with torch.random.fork_rng(devices=[]):          # the caller's global RNG stream is left alone
            torch.manual_seed(seed)
            self.discriminator = MomentumFeatureDiscriminator(n_domains, enc_c).to(dev)
        self.discriminator.synchronize_parameters()
        self.dis_optimizer = torch.optim.Adam(self.discriminator.parameters(), lr=lr)
        self.dis_criterion = CrossEntropy()
        # the discriminator's gradients live in one flat buffer: one all-reduce instead of one per parameter
        dparams = [p for p in self.discriminator.parameters() if p.requires_grad]
        self._dis_flat = torch.zeros(sum(p.numel() for p in dparams), dtype=torch.float32, device=dev)
        off = 0
        for p in dparams:
            p.grad = self._dis_flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.rewards = torch.zeros(M, dtype=torch.float32, device=dev)
        self.seed, self.epoch, self.step_idx = seed, 0, 0
        self.crop, self.scale_range = crop, scale_range
        use_dist = dist.is_available() and dist.is_initialized() if distributed is None else bool(distributed)
        self.world = dist.get_world_size() if use_dist else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.n_sources_total, self.src_offset = n_sources_total, src_offset
        self.policies = None
        self.searching = False           # becomes True at the warm-up -> search transition (begin_search)
        self.pretrained_steps = 0
        self.total_steps = 0
        self.use_graph = bool(graph)
        self._graphs = {}                # (kind, input shape) -> (GraphedTrainStep, images, labels)
        self._graph_pool = None
        self._reducer = BucketedAllReduce(model.store.grads) if self.world > 1 else None
        self._max_cloud = {}             # tuple(src_domains) -> largest per-domain source count of the global batch
        model.store.set_hyper(lr, weight_decay=weight_decay, grad_scale=1.0 / self.world)

    def close(self):
        """release the captured CUDA graphs (they hold NCCL kernels of the process group: drop them before
        torch.distributed.destroy_process_group())"""
        self._graphs.clear()
        self._graph_pool = None
        if self._reducer is not None:
            self._reducer.works = []

    # ---- epoch-level protocol (search_dg.py:323-347) -----------------------------------------------------------
    def set_lr(self, lr):
        self.lr = lr
        self.model.store.set_hyper(lr, weight_decay=self.wd, grad_scale=1.0 / self.world)

    def begin_search(self, lr_gamma=0.1):
        """The warm-up -> search transition of the reference: `discriminator.synchronize_parameters()` when
        epoch == WARMUP_EPOCH (search_dg.py:336-337) -- the EMA twin, whose hidden layer is the Sinkhorn point cloud,
        restarts from the warmed-up live discriminator -- and the model's MultiStepLR milestone at WARMUP_EPOCH
        (scheduler.py:11, gamma 0.1; the discriminator's schedule has gamma 1, scheduler.py:35).  Called automatically
        by the first set_policies() that follows pretrain_step()s; idempotent.

        Deliberate difference: the reference's synchronize_parameters() ALIASES the EMA parameters to the live ones
        (`param_k.data = param_q.data`), so during the first search epoch its "momentum" branch tracks the live
        discriminator and only detaches at the first momentum_update(); here the twin is a copy taken at the transition."""
        if self.searching:
            return
        self.discriminator.synchronize_parameters()
        if self.pretrained_steps > 0:
            self.set_lr(self.lr * lr_gamma)
        self.searching = True

    def set_policies(self, parsed_policies, epoch=None):
        """install the epoch's policies (search_dg.py:339-341) and reset the reward accumulator."""
        self.begin_search()
        self.policies = parsed_policies
        assert len(parsed_policies) == self.M
        if epoch is not None:
            self.epoch = epoch
        self.step_idx = 0
        self.rewards.zero_()

    def normalized_rewards(self):
        """search_dg.py:214 (raises if a reward is not finite)"""
        return SK.normalize_rewards(self.rewards)

    def end_epoch(self):
        self.discriminator.momentum_update()      # search_dg.py:346

    # ---- decisions ---------------------------------------------------------------------------------------------------
    def _layout(self, n_src):
        total = self.n_sources_total if self.n_sources_total is not None else n_src * self.world
        off = self.src_offset if self.src_offset is not None else n_src * self.rank
        return off, total

    def decision_rows(self, n_src, width, height, policies=None):
        off, total = self._layout(n_src)
        rows, raws = D.philox_rows(self.policies if policies is None else policies, n_src, width, height,
                                   self.crop or width, self.scale_range, self.seed, self.epoch, self.step_idx,
                                   scale_crop=self.crop is not None, src_offset=off, n_src_total=total)
        return rows, raws

    def domain_codes(self, src_domains, repeat):
        """ToTensor's random soft domain label (data/transform.py:260-274), one per source image,
        repeated for its M copies (data/transform.py:234)."""
        off, _ = self._layout(len(src_domains))
        dc = D.philox_soft_labels(src_domains, self.n_domains, self.seed, self.epoch, self.step_idx, off)
        return (np.repeat(dc, repeat, axis=0) if repeat > 1 else dc).astype(np.float32)

    def _global_max_cloud(self, src_domains):
        """largest number of source images of one domain in the GLOBAL batch = the size of the largest Sinkhorn cloud
        (every source contributes one row per policy); also checks that no domain is missing."""
        key = tuple(int(d) for d in src_domains)
        if key not in self._max_cloud:
            counts = torch.bincount(torch.tensor(key), minlength=self.n_domains).to(self.model.device)
            if self.world > 1:
                dist.all_reduce(counts)
            counts = counts.cpu()
            if int(counts.min()) == 0:
                raise RuntimeError("the global batch has no source image of domain %d: the diversity reward "
                                   "(search_dg.py:150-162) needs every domain in every step" % int(counts.argmin()))
            self._max_cloud[key] = int(counts.max())
        return self._max_cloud[key]

    # ---- the model part of a step ----------------------------------------------------------------------------------------
    def _augment(self, src_images, src_masks, rows, kind):
        """-> (images float32 [n,3,h,w], labels float32 [n,C,h,w]); written into the static buffers of the CUDA
        graph for this shape when one exists / is about to be captured"""
        n = len(rows)
        hw = self.crop if self.crop is not None else src_images.shape[1]
        ww = self.crop if self.crop is not None else src_images.shape[2]
        c = 2 if U8.DATASETS[self.dataset] == 0 else 1
        key = (kind, n, hw, ww)
        bufs = None
        if self.use_graph:
            ent = self._graphs.get(key)
            if ent is None:
                dev = src_images.device
                ent = [None, torch.empty((n, 3, hw, ww), dtype=torch.float32, device=dev),
                       torch.empty((n, c, hw, ww), dtype=torch.float32, device=dev)]
                self._graphs[key] = ent
            bufs = (ent[1], ent[2])
        if self.crop is not None:      # DGMultiPolicy -> DGRandomScaleCrop -> Normalize_dg -> ToTensor
            images, labels = U8.policy_scale_crop_normalize(src_images, src_masks, rows, self.crop, self.dataset,
                                                            out_images=bufs and bufs[0], out_labels=bufs and bufs[1])
        else:                          # DGMultiPolicy -> Normalize_dg -> ToTensor (one fused pass)
            images, labels = U8.policy_normalize(src_images, src_masks, rows, dataset=self.dataset,
                                                 out_images=bufs and bufs[0], out_labels=bufs and bufs[1])
        return images, labels, key

    def _model_step(self, images, labels, key):
        model, red = self.model, self._reducer
        if self.use_graph and self.total_steps >= 1:     # the very first step runs eagerly (lazy set-up)
            ent = self._graphs[key]
            if ent[0] is None:
                ent[0] = NW.GraphedTrainStep(model, images, labels, on_ready=red.ready if red else None,
                                             before_update=red.wait if red else None, pool=self._graph_pool)
                if self._graph_pool is None:
                    self._graph_pool = ent[0].graph.pool()
            return ent[0].replay()
        model.store.zero_grad()
        out = model.loss_step(images, labels, on_ready=red.ready if red else None)
        if red:
            red.wait()
        model.store.adam_step_dev()
        return out

    def _dis_step(self, dis_loss):
        self._dis_flat.zero_()
        dis_loss.backward()
        if self.world > 1:
            average_(self._dis_flat)
        self.dis_optimizer.step()

    def step(self, src_images, src_masks, src_domains, rows=None, dc=None, src_index=None):
        """src_images uint8 [S,H,W,3] (CUDA), src_masks uint8 [S,H,W], src_domains int [S] (host).
        src_index (int [S], host): src_images / src_masks are a whole RESIDENT POOL and the step's S sources are its
        entries src_index[0..S) (data.pool.ResidentPools.flat_indices) -- the augmentation kernels read them in place,
        no gather, no copy.  Returns dict(seg_loss, dis_loss, dice [classes], n_images) of 0-d / small CUDA tensors."""
        if not self.searching:
            self.begin_search()
        _, h, w, _ = src_images.shape
        s = len(src_domains)
        if src_index is None and src_images.shape[0] != s:
            raise ValueError("%d source images but %d domain codes" % (src_images.shape[0], s))
        if rows is None:
            rows, _ = self.decision_rows(s, w, h)
        if src_index is not None:
            src_index = np.asarray(src_index, np.int64)
            if src_index.shape != (s,) or src_index.min() < 0 or src_index.max() >= src_images.shape[0]:
                raise IndexError("src_index must hold %d indices into the %d pool images" % (s, src_images.shape[0]))
            rows = rows.copy()
            rows["src"] = src_index[rows["src"]]
        if dc is None:
            dc = self.domain_codes(src_domains, self.M)
        max_cloud = self._global_max_cloud(src_domains)
        dc_dev = torch.from_numpy(dc).to(src_images.device, non_blocking=True)
        images, labels, key = self._augment(src_images, src_masks, rows, "search")
        out = self._model_step(images, labels, key)
        feature = out["pooled"]
        # discriminator (search_dg.py:134-138)
        dis = self.discriminator
        _, domain_feature = dis(feature, momentum=True, return_feature=True)
        dis_loss = self.dis_criterion(dis(feature, momentum=False), dc_dev)
        # diversity reward over ALL ranks' features (search_dg.py:150-162)
        if self.world > 1:
            domain_feature, dc_all = gather_rows(domain_feature, dc_dev)
        else:
            dc_all = dc_dev
        SK.diversity_rewards(domain_feature, dc_all, self.M, self.rewards, max_cloud=max_cloud)
        self._dis_step(dis_loss)
        self.step_idx += 1
        self.total_steps += 1
        return dict(seg_loss=out["loss"], dis_loss=dis_loss.detach(), dice=dice_from_counts(out["counts"]),
                    n_images=images.shape[0])

    def pretrain_step(self, src_images, src_masks, src_domains, src_index=None):
        """One step of the warm-up phase (`pretrain`, search_dg.py:22-100): the un-augmented `sample['image']` (only
        DGRandomScaleCrop / Normalize_dg / ToTensor), segmentation step, live discriminator step on the detached pooled
        features; no policies, no rewards.  Same argument / return shapes as `step`."""
        _, h, w, _ = src_images.shape
        s = len(src_domains)
        _, raws = self.decision_rows(s, w, h, policies=[[[]]])
        if src_index is not None:          # the sources are entries of a resident pool (see step)
            raws = raws.copy()
            raws["src"] = np.asarray(src_index, np.int64)[raws["src"]]
        dc = self.domain_codes(src_domains, 1)
        dc_dev = torch.from_numpy(dc).to(src_images.device, non_blocking=True)
        images, labels, key = self._augment(src_images, src_masks, raws, "pretrain")
        out = self._model_step(images, labels, key)
        dis_loss = self.dis_criterion(self.discriminator(out["pooled"].detach(), momentum=False), dc_dev)
        self._dis_step(dis_loss)
        self.step_idx += 1
        self.total_steps += 1
        self.pretrained_steps += 1
        return dict(seg_loss=out["loss"], dis_loss=dis_loss.detach(), dice=dice_from_counts(out["counts"]),
                    n_images=images.shape[0])
