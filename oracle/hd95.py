"""ORACLE (test infrastructure, never on the product path).

Restatement of `medpy.metric.binary.hd95` / `__surface_distances` (MedPy 0.4.0, requirements.txt of the reference;
medpy itself is not installed here and not vendored: **parity unpinned** against medpy, pinned against the scipy /
numpy routines medpy is made of, which ARE installed) as validate() uses it, search_dg.py:246-260:

    result_border    = result ^ binary_erosion(result, structure=generate_binary_structure(ndim, 1), iterations=1)
    reference_border = reference ^ binary_erosion(reference, ...)
    dt  = distance_transform_edt(~reference_border, sampling=None)
    sds = dt[result_border]
    hd95 = numpy.percentile(numpy.hstack((sds(result, reference), sds(reference, result))), 95)
"""
import numpy as np
from scipy.ndimage import binary_erosion, distance_transform_edt, generate_binary_structure


def surface_distances(result, reference, connectivity=1):
    result = np.atleast_1d(result.astype(bool))
    reference = np.atleast_1d(reference.astype(bool))
    footprint = generate_binary_structure(result.ndim, connectivity)
    if 0 == np.count_nonzero(result):
        raise RuntimeError("The first supplied array does not contain any binary object.")
    if 0 == np.count_nonzero(reference):
        raise RuntimeError("The second supplied array does not contain any binary object.")
    result_border = result ^ binary_erosion(result, structure=footprint, iterations=1)
    reference_border = reference ^ binary_erosion(reference, structure=footprint, iterations=1)
    dt = distance_transform_edt(~reference_border, sampling=None)
    return dt[result_border]


def hd95(result, reference, percentile=95):
    hd1 = surface_distances(result, reference)
    hd2 = surface_distances(reference, result)
    return np.percentile(np.hstack((hd1, hd2)), percentile)


def brute_force(result, reference, percentile=95):
    """independent check for small masks: surfaces by explicit neighbour tests, distances by exhaustive search."""
    def surface(m):
        m = m.astype(bool)
        p = np.pad(m, 1)
        inner = p[1:-1, 1:-1] & p[:-2, 1:-1] & p[2:, 1:-1] & p[1:-1, :-2] & p[1:-1, 2:]
        return np.argwhere(m & ~inner)
    a, b = surface(result), surface(reference)

    def directed(p, q):
        d2 = ((p[:, None, :] - q[None, :, :]) ** 2).sum(-1)
        return np.sqrt(d2.min(1).astype(np.float64))
    return np.percentile(np.hstack((directed(a, b), directed(b, a))), percentile)


def random_blobs(rng, h, w, k):
    """test masks: the union of k random ellipses"""
    yy, xx = np.mgrid[0:h, 0:w]
    m = np.zeros((h, w), bool)
    for _ in range(k):
        cy, cx, ry, rx = rng.uniform(0, h), rng.uniform(0, w), rng.uniform(2, h / 3), rng.uniform(2, w / 3)
        m |= ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1
    return m
