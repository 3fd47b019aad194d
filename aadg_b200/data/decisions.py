"""Decision table: every random draw of the augmentation path made explicit.

The reference is unseeded and draws from Python `random` / NumPy global state inside DataLoader
workers.  Here each output image gets one fixed-size row
(`aadg_aug_row_t`, include/aadg_b200.h) holding *resolved integer/float parameters*; the CUDA bank
is a pure function of (source images, rows).  Two generators emit rows:

* `replay_sample`  — consumes `random.Random` / `np.random.RandomState` in exactly the reference's
  order (data/policy.py:15-30, data/basic.py:14,151-152, data/transform.py:35-55,104-135) so that a
  seeded reference run and this path see identical decisions (parity tests, golden files);
* `philox_rows`    — production: counter-based, keyed (seed, epoch, step, row), no host RNG state.
"""
import math
import numpy as np

from .basic import OP_ID, GEOMETRIC, MAX_OPS, level_to_value

ROW_DTYPE = np.dtype([
    ("src", "<i4"), ("n_ops", "<i4"), ("op", "<i4", (MAX_OPS,)), ("fparam", "<f4", (MAX_OPS,)),
    ("iparam", "<i4", (MAX_OPS, 6)),
    ("do_scale", "<i4"), ("scale_w", "<i4"), ("scale_h", "<i4"),
    ("pad", "<i4"), ("crop_x", "<i4"), ("crop_y", "<i4"),
])
assert ROW_DTYPE.itemsize == 160


def _fix(v):
    return int(math.floor(v * 65536.0 + 0.5))


def _affine_fixed(a):
    return (_fix(a[0]), _fix(a[1]), _fix(a[2] + a[0] * 0.5 + a[1] * 0.5),
            _fix(a[3]), _fix(a[4]), _fix(a[5] + a[3] * 0.5 + a[4] * 0.5))


def _geometric_coeffs(name, v, width, height):
    if name == "ShearX":
        return (1.0, v, 0.0, 0.0, 1.0, 0.0)
    if name == "ShearY":
        return (1.0, 0.0, 0.0, v, 1.0, 0.0)
    if name == "TranslateX":
        return (1.0, 0.0, v * width, 0.0, 1.0, 0.0)
    if name == "TranslateY":
        return (1.0, 0.0, 0.0, 0.0, 1.0, v * height)
    # Rotate: Pillow Image.rotate builds the inverse matrix about the image centre
    angle = v % 360.0
    if angle == 0:
        return (1.0, 0.0, 0.0, 0.0, 1.0, 0.0)
    ang = -math.radians(angle)
    m = [round(math.cos(ang), 15), round(math.sin(ang), 15), 0.0,
         round(-math.sin(ang), 15), round(math.cos(ang), 15), 0.0]
    cx, cy = width / 2.0, height / 2.0
    m2 = m[0] * -cx + m[1] * -cy + m[2]
    m5 = m[3] * -cx + m[4] * -cy + m[5]
    m[2], m[5] = m2 + cx, m5 + cy
    return tuple(m)


def resolve_op(row, k, name, level, width, height, ux=None, uy=None, mirror=False):
    """Fill slot k of `row` with op `name` at `level`, given the draws the reference would consume."""
    v = level_to_value(name, level)
    row["op"][k] = OP_ID[name]
    if name == "Solarize":
        row["iparam"][k, 0] = int(math.ceil(v))          # i < v  <=>  i < ceil(v) for integer i
    elif name == "Posterize":
        row["iparam"][k, 0] = (~(2 ** (8 - int(v)) - 1)) & 0xFF
    elif name in ("Contrast", "Color", "Brightness", "Sharpness"):
        row["fparam"][k] = np.float32(v)                  # Pillow casts the factor to a C float
    elif name == "Cutout":
        rect = (1, 1, 0, 0)                               # empty
        if v > 0.0:
            side = v * width
            x0 = int(max(0, ux - side / 2.0))
            y0 = int(max(0, uy - side / 2.0))
            rect = (x0, y0, int(min(width, x0 + side)), int(min(height, y0 + side)))
        row["iparam"][k, :4] = rect
    elif name in GEOMETRIC:
        if mirror:
            v = -v
        row["iparam"][k, :] = _affine_fixed(_geometric_coeffs(name, v, width, height))


def crop_padding(w, h, tw, th, padding=0):
    if padding > 0 or w < tw or h < th:
        return int(max(padding, max((tw - w) // 2 + 5, (th - h) // 2 + 5)))
    return 0


def _draw_scale_crop(py, width, height, tw, th, scale_range):
    """DGRandomScaleCrop.scale then RandomCrop (data/transform.py:104-112,35-55) for one image."""
    do_scale, w, h = 0, width, height
    if py.random() > 0.2:
        w = int(py.uniform(scale_range[0], scale_range[1]) * width)
        h = int(py.uniform(scale_range[0], scale_range[1]) * height)
        do_scale = 1
    pad = crop_padding(w, h, tw, th)
    w2, h2 = w + 2 * pad, h + 2 * pad
    x1 = y1 = 0
    if not (w2 == tw and h2 == th):
        x1 = py.randint(0, w2 - tw)
        y1 = py.randint(0, h2 - th)
    return do_scale, w, h, pad, x1, y1


class PolicyState:
    """Per-Policy call counter standing in for the reference's CutMix queue (data/policy.py:16-21):
    the queue only matters through the `random.choice(queue)` draw it triggers while len <= 10."""

    def __init__(self, m):
        self.calls = [0] * m


def replay_policy_call(row, policy, n_calls, src_index, width, height, py, npr):
    """One `Policy.__call__` (data/policy.py:15-30) resolved into `row`: the CutMix-queue draw (only while the queue
    holds <= 10 entries; `n_calls` = calls of this Policy object so far, this one included), the sub-policy choice and
    the per-op draws, consumed from `py` (random.Random or the `random` module) and `npr` (RandomState or
    `np.random`) in the reference's order."""
    qlen = min(n_calls, 11)
    if qlen <= 10:
        py.choice(range(qlen))                        # pair = random.choice(self.queue)
    sub = py.choice(policy)
    row["src"] = src_index
    row["n_ops"] = len(sub)
    for k, (name, level) in enumerate(sub):
        ux = uy = None
        mirror = False
        if name == "Cutout" and level_to_value(name, level) > 0.0:
            ux = npr.uniform(width)
            uy = npr.uniform(height)
        elif name in GEOMETRIC:
            mirror = py.random() > 0.5
        resolve_op(row, k, name, level, width, height, ux, uy, mirror)


def replay_sample(parsed_policies, src_index, width, height, crop, scale_range, py, npr, state,
                  scale_crop=True):
    """Rows for the M augmented copies of one source image, consuming `py` (random.Random) and
    `npr` (np.random.RandomState) exactly like DGMultiPolicy -> DGRandomScaleCrop do.

    Returns (rows[M], raw_row) — raw_row carries the scale/crop decision of the un-augmented
    'image' entry, which the reference draws *before* the copies (data/transform.py:123-124)."""
    m = len(parsed_policies)
    rows = np.zeros(m, ROW_DTYPE)
    tw = th = crop
    for j, policy in enumerate(parsed_policies):
        state.calls[j] += 1
        replay_policy_call(rows[j], policy, state.calls[j], src_index, width, height, py, npr)
    raw = np.zeros(1, ROW_DTYPE)[0]
    raw["src"] = src_index
    if scale_crop:
        (raw["do_scale"], raw["scale_w"], raw["scale_h"], raw["pad"], raw["crop_x"],
         raw["crop_y"]) = _draw_scale_crop(py, width, height, tw, th, scale_range)
        for j in range(m):
            r = rows[j]
            (r["do_scale"], r["scale_w"], r["scale_h"], r["pad"], r["crop_x"],
             r["crop_y"]) = _draw_scale_crop(py, width, height, tw, th, scale_range)
    else:
        for r in list(rows) + [raw]:
            r["scale_w"], r["scale_h"] = width, height
    return rows, raw


def soft_label(py, domain, n_domains):
    """ToTensor's SoftLable(ToMultiLabel(dc, n)) (data/transform.py:251-274) with explicit RNG."""
    new = [0.0] * n_domains
    new[domain] = 0.8 + py.random() * 0.2
    acc = new[domain]
    for i in range(n_domains):
        if i != domain:
            if i == n_domains - 1:
                new[i] = 1 - acc
            else:
                new[i] = py.random() * (1 - acc)
                acc += new[i]
    return np.asarray(new, np.float32)


class _UniformStream:
    """`.random()` over a fixed array of uniforms (lets soft_label consume counter-based draws)"""

    def __init__(self, values):
        self.values, self.i = values, 0

    def random(self):
        v = float(self.values[self.i])
        self.i += 1
        return v


def philox_soft_labels(domains, n_domains, seed, epoch=0, step=0, src_offset=0):
    """ToTensor's random soft domain labels (data/transform.py:260-274) for source images src_offset.. of the global
    batch, counter-based: keyed (seed, epoch, step, global source index), independent of how the batch is sharded."""
    domains = [int(d) for d in domains]
    ids = src_offset + np.arange(len(domains))
    u = _uniforms((int(seed) ^ 0x50F7) & 0xFFFFFFFFFFFFFFFF, epoch, step, len(domains), (n_domains + 3) // 4, ids)
    return np.stack([soft_label(_UniformStream(u[i]), d, n_domains) for i, d in enumerate(domains)])


# ------------------------------------------------------------------------------------------------
# production generator: Philox4x32-10, counter = (row, draw, step, epoch), key = seed
# ------------------------------------------------------------------------------------------------
_PH_M0, _PH_M1 = 0xD2511F53, 0xCD9E8D57
_PH_W0, _PH_W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32(counter, key):
    """Philox4x32-10 on uint32 arrays: counter [...,4], key [...,2] -> [...,4] uint32."""
    c = [np.asarray(counter[..., i], np.uint64) for i in range(4)]
    k0 = np.asarray(key[..., 0], np.uint64)
    k1 = np.asarray(key[..., 1], np.uint64)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(_PH_M0) * c[0]
        p1 = np.uint64(_PH_M1) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = [(hi1 ^ c[1] ^ k0) & mask, lo1, (hi0 ^ c[3] ^ k1) & mask, lo0]
        k0 = (k0 + np.uint64(_PH_W0)) & mask
        k1 = (k1 + np.uint64(_PH_W1)) & mask
    return np.stack(c, axis=-1).astype(np.uint32)


def _uniforms(seed, epoch, step, n_rows, n_draws, row_ids=None):
    """float64 uniforms in [0,1) with 32 bits of entropy, shape [n_rows, n_draws*4]; row_ids = the counter value of
    every row (default 0..n_rows-1)."""
    ctr = np.zeros((n_rows, n_draws, 4), np.uint32)
    ids = np.arange(n_rows, dtype=np.uint32) if row_ids is None else np.asarray(row_ids, np.uint32)
    ctr[..., 0] = ids[:, None]
    ctr[..., 1] = np.arange(n_draws, dtype=np.uint32)[None, :]
    ctr[..., 2] = np.uint32(step & 0xFFFFFFFF)
    ctr[..., 3] = np.uint32(epoch & 0xFFFFFFFF)
    key = np.zeros((n_rows, n_draws, 2), np.uint32)
    key[..., 0] = np.uint32(seed & 0xFFFFFFFF)
    key[..., 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    bits = philox4x32(ctr, key).reshape(n_rows, n_draws * 4)
    return bits.astype(np.float64) / 4294967296.0


def philox_rows(parsed_policies, n_src, width, height, crop, scale_range, seed, epoch=0, step=0,
                scale_crop=True, src_offset=0, n_src_total=None):
    """Rows for n_src source images x M policies, row index = s*M + j (the reference's collate
    order, data/transform.py:323-340).  Same distributions as the reference's draws, different
    (counter-based) stream.  Also returns the raw-image rows [n_src].

    src_offset / n_src_total: these n_src images are sources src_offset .. src_offset+n_src-1 of a GLOBAL batch of
    n_src_total; the draws are keyed by the global row index, so a rank that owns a shard of the batch makes exactly the
    decisions a single process holding the whole batch makes for those images (1-vs-N result parity)."""
    m = len(parsed_policies)
    n = n_src * m
    total = n_src if n_src_total is None else int(n_src_total)
    ids = np.concatenate([np.arange(src_offset * m, (src_offset + n_src) * m),
                          total * m + np.arange(src_offset, src_offset + n_src)])
    u = _uniforms(seed, epoch, step, n + n_src, 5, ids)   # 20 uniforms per row
    rows = np.zeros(n, ROW_DTYPE)
    raws = np.zeros(n_src, ROW_DTYPE)
    tw = th = crop

    def scale_crop_from(ur, r):
        w, h = width, height
        if ur[13] > 0.2:
            w = int((scale_range[0] + (scale_range[1] - scale_range[0]) * ur[14]) * width)
            h = int((scale_range[0] + (scale_range[1] - scale_range[0]) * ur[15]) * height)
            r["do_scale"] = 1
        pad = crop_padding(w, h, tw, th)
        r["scale_w"], r["scale_h"], r["pad"] = w, h, pad
        w2, h2 = w + 2 * pad, h + 2 * pad
        if not (w2 == tw and h2 == th):
            r["crop_x"] = int(ur[16] * (w2 - tw + 1))
            r["crop_y"] = int(ur[17] * (h2 - th + 1))

    for s in range(n_src):
        for j, policy in enumerate(parsed_policies):
            i = s * m + j
            ur = u[i]
            row = rows[i]
            sub = policy[int(ur[0] * len(policy))]
            row["src"] = s
            row["n_ops"] = len(sub)
            for k, (name, level) in enumerate(sub):
                # Cutout centre ~ np.random.uniform(w) == w + (1-w)*u  (data/basic.py:151-152)
                ux = width + (1.0 - width) * ur[1 + 3 * k]
                uy = height + (1.0 - height) * ur[2 + 3 * k]
                mirror = ur[3 + 3 * k] > 0.5
                resolve_op(row, k, name, level, width, height, ux, uy, mirror)
            if scale_crop:
                scale_crop_from(ur, row)
            else:
                row["scale_w"], row["scale_h"] = width, height
        raws[s]["src"] = s
        if scale_crop:
            scale_crop_from(u[n + s], raws[s])
        else:
            raws[s]["scale_w"], raws[s]["scale_h"] = width, height
    return rows, raws
