"""ORACLE (test infrastructure).  The controller's counter-based categorical sampling restated in numpy: decision t of
batch row m draws u = (Philox4x32-10(key = seed, counter = (m, t, call lo, call hi))[0] >> 8) * 2**-24 and takes the
first k with cumsum(p)[k] > u (float32 running sum, last index if none) -- csrc/controller.cu walk_kernel.  The
probabilities themselves (LSTM + heads) are checked against the torch Controller mirror of models/controller.py."""
import numpy as np

from aadg_b200.data.decisions import philox4x32


def sample_actions(step_probs, n_ops, n_mags, seed, call):
    """step_probs float32 [M, steps, Vmax] (the kernel's own probabilities) -> int64 [M, steps]"""
    m, steps, _ = step_probs.shape
    out = np.zeros((m, steps), np.int64)
    for i in range(m):
        for t in range(steps):
            ctr = np.array([i, t, call & 0xFFFFFFFF, call >> 32], np.uint32)
            key = np.array([seed & 0xFFFFFFFF, seed >> 32], np.uint32)
            r = philox4x32(ctr, key)
            u = np.float32(int(r[0]) >> 8) * np.float32(1.0 / 16777216.0)
            v = n_mags if t & 1 else n_ops
            cum = np.float32(0)
            a = v - 1
            for k in range(v):
                cum = np.float32(cum + step_probs[i, t, k])
                if cum > u:
                    a = k
                    break
            out[i, t] = a
    return out
