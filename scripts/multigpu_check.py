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
    eng.close()
    if rank == 0:
        print("MULTIGPU_CHECK LOCKSTEP%s" % ("_GRAPH" if graph else ""), "OK" if ok else "FAILED", "world", world,
              "rewards", eng.rewards.cpu().numpy().round(5).tolist(), "param checksums", all_sums.cpu().numpy().tolist(),
              "| gradient buckets handed to NCCL per step (elements):", buckets[:len(buckets) // 3 or None])
    return ok


def parity(rank, world, dev):
    """The same global batch on N ranks (SyncBN statistics) and on one GPU: loss, rewards, parameters.

    The engine is not bitwise reproducible: its batch-norm statistics and weight gradients are summed with fp32 atomics,
    and bf16 storage re-quantises any last-bit difference up to the bf16 noise floor within a few layers
    (tests/test_parity_gpu.py, DESIGN.md "Precision").  "Equal" therefore means: the N-GPU run differs from the 1-GPU run
    by no more than two 1-GPU runs of the same batch differ from each other (x4 margin; small absolute floors)."""
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
        losses, rewards = [], []
        for _ in range(2):
            out = eng.step(x, m, [domains[i] for i in idx])
            losses.append(out["seg_loss"].reshape(1).clone())
            rewards.append(eng.rewards.clone())
        NW.set_sync_bn(None)
        return torch.cat(losses), torch.stack(rewards), model.store.params.clone()
    l_n, r_n, p_n = run(True, mine)
    dist.all_reduce(l_n)
    l_n /= world                                                       # global loss = mean of the equal-sized shards' losses
    ok, msg = True, ""
    if rank == 0:
        l_a, r_a, p_a = run(False, list(range(n_src)))
        l_b, r_b, p_b = run(False, list(range(n_src)))                 # the 1-GPU run again: the engine's own spread
        rel = lambda u, v: ((u - v).abs() / v.abs()).cpu().numpy()      # noqa: E731
        e_loss, s_loss = rel(l_n, l_a), rel(l_b, l_a)
        e_rew, s_rew = rel(r_n, r_a).max(axis=1), rel(r_b, r_a).max(axis=1)
        e_par = ((p_n - p_a).norm() / p_a.norm()).item()
        s_par = ((p_b - p_a).norm() / p_a.norm()).item()
        ok = bool((e_loss <= 4 * s_loss + 1e-4).all() and (e_rew <= 4 * s_rew + 2e-2).all() and e_par <= 4 * s_par + 1e-3)
        msg = ("loss per step 1-GPU %s N-GPU %s | rel diff N-vs-1 %s, 1-GPU run-to-run %s | rewards (cumulative, per step) "
               "N-vs-1 %s, run-to-run %s | params rel L2 N-vs-1 %.2e, run-to-run %.2e" %
               (l_a.cpu().numpy().tolist(), l_n.cpu().numpy().tolist(), e_loss.tolist(), s_loss.tolist(), e_rew.tolist(),
                s_rew.tolist(), e_par, s_par))
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
    from aadg_b200.host.search import shutdown
    sys.stdout.flush()
    if not all(results):
        sys.stderr.write("MULTIGPU_CHECK: a check failed\n")
    rc = 0 if all(results) else 1
    import threading
    threading.Timer(45.0, lambda: os._exit(rc)).start()      # teardown watchdog (see search.shutdown)
    shutdown()
    os._exit(rc)


if __name__ == "__main__":
    main()
