"""GPU property tests (hypothesis): the CUDA uint8 bank against the oracle over op x magnitude x image
content x size, and N4's resident-pool gather on the device.  Bit exact."""
import numpy as np
import pytest
import torch

hyp = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st, HealthCheck  # noqa: E402

from aadg_b200.data import decisions as D  # noqa: E402
from aadg_b200.data.basic import AADG_OPS  # noqa: E402
from oracle import u8_policy as P  # noqa: E402
from test_property_cpu import images  # noqa: E402

pytestmark = pytest.mark.gpu
NAMES = [n for n, _, _ in AADG_OPS]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@st.composite
def chains(draw):
    n = draw(st.integers(1, 4))
    return [(draw(st.sampled_from(NAMES)), draw(st.integers(0, 9)) / 9) for _ in range(n)]


@settings(max_examples=120, deadline=None, suppress_health_check=list(HealthCheck))
@given(img=images(8, 72), chain_list=st.lists(chains(), min_size=1, max_size=6), seed=st.integers(0, 2 ** 31 - 1))
def test_bank_equals_oracle_on_generated_chains(img, chain_list, seed):
    from aadg_b200.ops import u8
    from test_u8_gpu import make_row
    h, w = img.shape[:2]
    rng = np.random.RandomState(seed)
    mask = (rng.randint(0, 3, (h, w)) * 127).astype(np.uint8)
    rows = np.stack([make_row(0, c, w, h, rng) for c in chain_list])
    out, outm = u8.apply_policy(dev(img[None]), dev(mask[None]), rows, want_masks=True)
    out, outm = out.cpu().numpy(), outm.cpu().numpy()
    for i, c in enumerate(chain_list):
        wi, wm = P.apply_chain(img, mask, rows[i])
        assert np.array_equal(out[i], wi), (c, img.shape, int(np.abs(out[i].astype(int) - wi).max()))
        assert np.array_equal(outm[i], wm), (c, img.shape)


@settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))
@given(img=images(16, 64), chain_list=st.lists(chains(), min_size=2, max_size=4), seed=st.integers(0, 2 ** 31 - 1),
       crop=st.integers(8, 48), dataset=st.sampled_from(["optic", "vessel"]))
def test_policy_scale_crop_normalize_equals_oracle(img, chain_list, seed, crop, dataset):
    """the whole train transform (policy -> DGRandomScaleCrop -> Normalize_dg -> ToTensor) on generated decisions"""
    from aadg_b200.ops import u8
    from test_u8_gpu import make_row
    h, w = img.shape[:2]
    rng = np.random.RandomState(seed)
    mask = (rng.randint(0, 3, (h, w)) * 127).astype(np.uint8)
    rows = np.stack([make_row(0, c, w, h, rng) for c in chain_list])
    for row in rows:
        row["do_scale"] = int(rng.rand() > 0.2)
        sw, sh = (int(rng.uniform(0.5, 2) * w), int(rng.uniform(0.5, 2) * h)) if row["do_scale"] else (w, h)
        sw, sh = max(sw, 1), max(sh, 1)
        row["scale_w"], row["scale_h"] = sw, sh
        pad = D.crop_padding(sw, sh, crop, crop)
        row["pad"] = pad
        row["crop_x"] = rng.randint(0, sw + 2 * pad - crop + 1)
        row["crop_y"] = rng.randint(0, sh + 2 * pad - crop + 1)
    im, lb = u8.policy_scale_crop_normalize(dev(img[None]), dev(mask[None]), rows, crop, dataset)
    want = P.apply_rows(img[None], mask[None], rows, crop=crop, dataset=dataset)
    assert np.array_equal(im.cpu().numpy(), want["images"]), (img.shape, crop)
    assert np.array_equal(lb.cpu().numpy(), want["labels"]), (img.shape, crop)


def test_resident_pools_gather_on_the_device():
    """N4 (data/optic.py:79-91, data/transform.py:323-340): pools resident in HBM, a step's sources are one gather in the
    reference's sampling order; the gathered batch feeds the bank directly"""
    from aadg_b200.data.pool import ResidentPools
    from aadg_b200.ops import u8
    rng = np.random.RandomState(0)
    sizes = {"DGS": 5, "RIM": 9, "REF": 3}
    imgs = {k: rng.randint(0, 256, (n, 32, 40, 3)).astype(np.uint8) for k, n in sizes.items()}
    msks = {k: (rng.randint(0, 3, (n, 32, 40)) * 127).astype(np.uint8) for k, n in sizes.items()}
    pools = ResidentPools(imgs, msks, device="cuda")
    np.random.seed(123)
    want = [[np.random.choice(n, 1)[0] for n in sizes.values()] for _ in range(4)]      # the reference's draws
    np.random.seed(123)
    idx = pools.sample_indices(4)
    assert idx.tolist() == want
    x, m, dom = pools.gather(idx)
    assert x.is_cuda and m.is_cuda and x.shape == (12, 32, 40, 3) and dom == [0, 1, 2] * 4
    keys = list(sizes)
    for b in range(4):
        for d in range(3):
            assert np.array_equal(x[b * 3 + d].cpu().numpy(), imgs[keys[d]][want[b][d]])
            assert np.array_equal(m[b * 3 + d].cpu().numpy(), msks[keys[d]][want[b][d]])
    fi, fl = u8.normalize_to_tensor(x, m, "optic")
    ref = (x.cpu().numpy().astype(np.float32) / np.float32(127.5) - np.float32(1)).transpose(0, 3, 1, 2)
    assert np.array_equal(fi.cpu().numpy(), ref) and fl.shape == (12, 2, 32, 40)
