"""Run under torchrun (one rank per GPU).  Three checks of the sharded search step:

  LOCKSTEP   every rank owns different source images; after two steps the parameters are bit-identical on all
             ranks (same averaged gradients), the rewards are identical everywhere, the loss is finite; run eagerly
             and with the step captured in a CUDA graph (bucketed all-reduce inside the graph).
  PARITY     1-vs-N on the SAME global batch: the ranks shard a global batch (SyncBN statistics on) and rank 0 then
             runs the whole batch alone on a fresh identical model; loss / rewards / parameters must agree
             (tolerances below; the residual is the summation order of fp32 atomics and of the all-reduce).
  OVERLAP    prints the bucket sizes the backward pass handed to NCCL.

Prints one line per check: `MULTIGPU_CHECK <name> OK|FAILED ...`; exit code 0 only if all pass."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class Cfg:
    class CONTROLLER:
        EXCLUDE_OPS = []
        L = 2
        NUM_MAGS = 10
        EXCLUDE_OPS_NUM = 0
    SEED = 0


def all_ok(ok, dev):
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return flag.item() == 1.0


def lockstep(rank, world, dev, graph):
    from aadg_b200.data.policy import parse_policies
    from aadg_b200.host.search import SearchEngine, gather_rows
    from aadg_b200.nn import DeepLabV3Plus
    from aadg_b200.synth import fundus_batch, random_policies
    model = DeepLabV3Plus(encoder_name="resnet18", encoder_weights=None, in_channels=3, classes=2,
                          aux_params=dict(pooling="avg"), device=dev, seed=5)
    eng = SearchEngine(model, n_domains=3, M=6, seed=5, graph=graph)
    eng.set_policies(parse_policies(random_policies(seed=5), Cfg), epoch=0)
    imgs, masks = fundus_batch(6, 128, 128, seed=100 + rank)          # each rank owns different source images
    d_imgs, d_masks = torch.from_numpy(imgs).to(dev), torch.from_numpy(masks).to(dev)
    domains = [i % 3 for i in range(6)]
    ok = True
    buckets = []
    if eng._reducer is not None:
        orig = eng._reducer.ready

        def spy(off):
            hi = eng._reducer.hi
            orig(off)
            if eng._reducer.hi != hi:
                buckets.append(hi - eng._reducer.hi)
        eng._reducer.ready = spy
    for step in range(3):
        out = eng.step(d_imgs, d_masks, domains)
        ok &= bool(torch.isfinite(out["seg_loss"]).all())
    sums = torch.stack([model.store.params.double().sum(), model.store.params.double().abs().sum(),
                        eng.rewards.double().sum()]).reshape(1, 3)
    (all_sums,) = gather_rows(sums.float())
    ok &= bool((all_sums == all_sums[0:1]).all())
    (all_rewards,) = gather_rows(eng.rewards.reshape(1, -1))
    ok &= bool((all_rewards == all_rewards[0:1]).all()) and bool((eng.rewards > 0).all())
    if graph:
        ok &= all(v[0] is not None for v in eng._graphs.values())
    ok = all_ok(ok, dev)
    if rank == 0:
        print("MULTIGPU_CHECK LOCKSTEP%s" % ("_GRAPH" if graph else ""), "OK" if ok else "FAILED", "world", world,
              "rewards", eng.rewards.cpu().numpy().round(5).tolist(), "param checksums", all_sums.cpu().numpy().tolist(),
              "| gradient buckets handed to NCCL per step (elements):", buckets[:len(buckets) // 3 or None])
    return ok


def parity(rank, world, dev):
    """same global batch on N ranks (SyncBN) and on one: loss, rewards, parameters"""
    from aadg_b200.data.policy import parse_policies
    from aadg_b200.host.search import SearchEngine, shard_sources
    from aadg_b200.nn import DeepLabV3Plus
    from aadg_b200.nn import network as NW
    from aadg_b200.synth import fundus_batch, random_policies
    n_src = 6 * world if (6 * world) % 3 == 0 else 12 * world
    imgs, masks = fundus_batch(n_src, 128, 128, seed=77)              # the same global batch on every rank
    domains = [i % 3 for i in range(n_src)]
    parsed = parse_policies(random_policies(seed=5), Cfg)
    mine = shard_sources(n_src, rank, world)

    def run(distributed, idx):
        NW.set_sync_bn(True if distributed else None)
        model = DeepLabV3Plus(encoder_name="resnet18", encoder_weights=None, in_channels=3, classes=2,
                              aux_params=dict(pooling="avg"), device=dev, seed=5)
        model.dropout_enabled = False      # the dropout mask is keyed by the LOCAL element index
        eng = SearchEngine(model, n_domains=3, M=6, seed=5, crop=128, distributed=distributed,
                           n_sources_total=n_src, src_offset=idx[0])
        eng.set_policies(parsed, epoch=0)
        x, m = torch.from_numpy(imgs[idx]).to(dev), torch.from_numpy(masks[idx]).to(dev)
        losses = []
        for _ in range(2):
            out = eng.step(x, m, [domains[i] for i in idx])
            losses.append(out["seg_loss"].reshape(1).clone())
        NW.set_sync_bn(None)
        return torch.cat(losses), eng.rewards.clone(), model.store.params.clone()
    l_n, r_n, p_n = run(True, mine)
    dist.all_reduce(l_n)
    l_n /= world                                                       # global loss = mean of the equal-sized shards' losses
    ok, msg = True, ""
    if rank == 0:
        l_1, r_1, p_1 = run(False, list(range(n_src)))
        e_loss = ((l_n - l_1).abs() / l_1.abs()).cpu().numpy()
        e_rew = ((r_n - r_1).abs() / r_1.abs()).max().item()
        e_par = ((p_n - p_1).norm() / p_1.norm()).item()
        # step 0: same weights, so only summation order differs (1e-5); step 1 follows an Adam step, whose
        # sign-like first update amplifies gradient noise on near-zero gradients (1e-3)
        ok = bool(e_loss[0] <= 1e-5 and e_loss[1] <= 1e-3 and e_rew <= 1e-3 and e_par <= 1e-3)
        msg = "loss 1-GPU %s N-GPU %s rel %s | rewards rel %.2e | params rel L2 %.2e" % (
            l_1.cpu().numpy().tolist(), l_n.cpu().numpy().tolist(), e_loss.tolist(), e_rew, e_par)
    ok = all_ok(ok, dev)
    if rank == 0:
        print("MULTIGPU_CHECK PARITY_1_vs_%d" % world, "OK" if ok else "FAILED", "global batch %d sources x 6 policies, "
              "SyncBN statistics:" % n_src, msg)
    return ok


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    results = [lockstep(rank, world, dev, False), lockstep(rank, world, dev, True), parity(rank, world, dev)]
    dist.destroy_process_group()
    sys.exit(0 if all(results) else 1)


if __name__ == "__main__":
    main()
