"""The numbers behind the TopK search's first probe (spgemm-new_b200/csrc/topk_select.cuh: TopkSeed, topk_seed()):
the expected RANKS of the order statistics it interpolates between are properties of the lane layout alone -- they
must come out the same for any continuous distribution of the values.  Pure numpy; no GPU, no library."""
import numpy as np
import pytest

# topk_seed(): ranks (of 256) of the minimum over the lane maxima (A), over the maxima of lane pairs (B), of lane
# quads (C) and over the maxima of the lanes' four-element groups (G)
RA, RB, RC, RG = 100.6, 48.3, 20.7, 175.0


def _stats(x):
    """x [rows, 256] -> (G, A, B, C, max) per row in the kernel's layout: lane l holds columns 4l..4l+3 and 128+4l.."""
    r = x.shape[0]
    lanes = np.concatenate([x[:, :128].reshape(r, 32, 4), x[:, 128:].reshape(r, 32, 4)], axis=2)   # [r, lane, 8]
    group_max = lanes.reshape(r, 64, 4).max(2)
    lane_max = lanes.max(2)
    pair_max = np.maximum(lane_max[:, 0::2], lane_max[:, 1::2])
    quad_max = np.maximum(pair_max[:, 0::2], pair_max[:, 1::2])
    return group_max.min(1), lane_max.min(1), pair_max.min(1), quad_max.min(1), lane_max.max(1)


@pytest.mark.parametrize("dist", ["uniform", "normal", "lognormal", "cauchy"])
def test_order_statistics_have_the_ranks_the_seed_assumes(dist):
    rng = np.random.default_rng(11)
    x = {"uniform": rng.random, "normal": rng.standard_normal,
         "lognormal": lambda s: np.exp(3 * rng.standard_normal(s)), "cauchy": rng.standard_cauchy}[dist]((6000, 256))
    g, a, b, c, top = _stats(x)
    rank = lambda v: (x >= v[:, None]).sum(1)
    for name, v, want in (("G", g, RG), ("A", a, RA), ("B", b, RB), ("C", c, RC)):
        assert abs(rank(v).mean() - want) < 1.5, (dist, name, rank(v).mean())
    # the bracket's lower ends are valid: at least 32 (one per lane) / 64 (one per group) elements are not below them
    assert rank(a).min() >= 32 and rank(g).min() >= 64 and rank(top).max() >= 1


@pytest.mark.parametrize("k", [1, 8, 16, 20, 21, 32, 48, 49, 64])
def test_first_threshold_lands_near_rank_k(k):
    """The interpolation BY RANK between the enclosing statistics: its count is within a few elements of k on average
    for a smooth distribution, and the distance does not depend on which one (what the min / max interpolation it
    replaced could not do: ~26 elements off at k = 32 on normal data)."""
    rng = np.random.default_rng(k)
    for dist in ("uniform", "normal"):
        x = (rng.random((3000, 256)) if dist == "uniform" else rng.standard_normal((3000, 256))).astype(np.float32)
        g, a, b, c, top = _stats(x)
        if k <= 20:
            lo, hi, w = c, top, (RC - k) / (RC - 1.0)
        elif k <= 48:
            lo, hi, w = b, c, (RB - k) / (RB - RC)
        else:
            lo, hi, w = a, b, (RA - k) / (RA - RB)
        t = lo + (hi - lo) * w
        off = np.abs((x >= t[:, None]).sum(1) - k)
        assert off.mean() < 13.0, (dist, k, off.mean())
    xn = rng.standard_normal((3000, 256)).astype(np.float32)
    lo, hi = xn.min(1), xn.max(1)
    t_old = lo + (hi - lo) * ((256 - 32) / (256 - 32 + 31.0))
    assert np.abs((xn >= t_old[:, None]).sum(1) - 32).mean() > 20.0   # the seed this replaces, k = 32, normal data
