"""Synthetic graphs with the shapes BASELINE.json names (no datasets: there is no network).

All generators are torch-only and run on CPU or GPU from a fixed seed:
    flickr    N=89 250     E~0.9 M (+self loops)   power-law, mean degree 10
    reddit    N=232 965    E=114 615 892           heavy tail, mean 492
    products  N=2 449 029  E~61.86 M               power-law, mean 25
    proteins  N=132 534    E~39.56 M               dense near-uniform, mean 298
    rmat      R-MAT (0.57, 0.19, 0.19, 0.05), 2^s nodes, given average degree
Destination rows get log-normal degrees rescaled to the target E; sources are drawn
proportionally to a log-normal popularity (hubs), sorted within each row (CSR order).
"""
from __future__ import annotations

import math
from typing import Dict

import torch

SHAPES = {
    #            N          E            kind        sigma  self_loops
    "flickr":   (89_250,    899_756,     "powerlaw", 1.0,   True),
    "reddit":   (232_965,   114_615_892, "powerlaw", 1.1,   False),
    "products": (2_449_029, 61_859_140,  "powerlaw", 1.3,   False),
    "proteins": (132_534,   39_561_252,  "uniform",  0.0,   False),
}


def _lognormal_degrees(n: int, e: int, sigma: float, gen, device) -> torch.Tensor:
    raw = torch.exp(sigma * torch.randn(n, generator=gen, device=device, dtype=torch.float64))
    deg = torch.floor(raw * (e / raw.sum())).to(torch.int64)
    short = e - int(deg.sum())
    if short > 0:  # hand the rounding remainder to random rows
        bump = torch.randint(0, n, (short,), generator=gen, device=device)
        deg += torch.bincount(bump, minlength=n)
    return deg


def csr_from_degrees(deg: torch.Tensor, n_cols: int, gen, kind: str = "powerlaw", sigma: float = 1.2):
    device = deg.device
    n = deg.numel()
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    torch.cumsum(deg, 0, out=indptr[1:])
    e = int(indptr[-1])
    assert e < 2**31, "int32 CSR offsets (reference format) need E < 2^31"
    if kind == "uniform":
        cols = torch.randint(0, n_cols, (e,), generator=gen, device=device, dtype=torch.int64)
    else:
        pop = torch.exp(sigma * torch.randn(n_cols, generator=gen, device=device, dtype=torch.float64))
        cdf = torch.cumsum(pop, 0)
        cols = torch.empty(e, dtype=torch.int64, device=device)
        step = 1 << 26
        for s in range(0, e, step):  # chunked to bound the fp64 scratch
            u = torch.rand(min(step, e - s), generator=gen, device=device, dtype=torch.float64) * cdf[-1]
            cols[s:s + u.numel()] = torch.searchsorted(cdf, u).clamp_(max=n_cols - 1)
    rows = torch.repeat_interleave(torch.arange(n, device=device), deg)
    key = rows * n_cols + cols
    key, _ = torch.sort(key)
    cols = (key - rows * n_cols)  # rows are already sorted, so `rows` is unchanged by the sort
    return indptr.to(torch.int32), cols.to(torch.int32)


def make_graph(shape: str, seed: int = 0, device="cpu", scale: float = 1.0,
               values: str = "uniform") -> Dict[str, torch.Tensor]:
    """values: "uniform" U[0,1) (main.cu input_mode 1), "ones", "mean" (1/deg), "gcn" (sym-norm)."""
    n, e, kind, sigma, self_loops = SHAPES[shape]
    n = max(8, int(round(n * scale)))
    e = max(n, int(round(e * scale)))
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    if kind == "uniform":
        mean = e / n
        deg = torch.randint(int(mean * 0.75), int(mean * 1.25) + 2, (n,), generator=gen, device=device)
    else:
        deg = _lognormal_degrees(n, e, sigma, gen, device)
    if self_loops:
        deg = deg + 1
    indptr, indices = csr_from_degrees(deg, n, gen, kind, sigma if sigma else 1.2)
    if self_loops:  # overwrite one slot per row with the diagonal (keeps E exact), re-sort not needed for math
        first = indptr[:-1].long()
        indices[first] = torch.arange(n, device=device, dtype=torch.int32)
    return finish_graph(indptr, indices, n, gen, values, shape)


def finish_graph(indptr, indices, n_cols, gen, values, name):
    device = indptr.device
    e = indices.numel()
    n = indptr.numel() - 1
    deg = (indptr[1:] - indptr[:-1]).to(torch.float32)
    if values == "uniform":
        val = torch.rand(e, generator=gen, device=device, dtype=torch.float32)
    elif values == "ones":
        val = torch.ones(e, device=device, dtype=torch.float32)
    elif values == "mean":
        val = torch.repeat_interleave(1.0 / deg.clamp(min=1), deg.long())
    elif values == "gcn":
        din = deg.clamp(min=1)
        dout = torch.bincount(indices.long(), minlength=n_cols).to(torch.float32).clamp(min=1)
        rows = torch.repeat_interleave(torch.arange(n, device=device), deg.long())
        val = 1.0 / torch.sqrt(din[rows] * dout[indices.long()])
    else:
        raise ValueError(values)
    return {"indptr": indptr.contiguous(), "indices": indices.contiguous(), "values": val.contiguous(),
            "v_num": n, "e_num": e, "n_cols": n_cols, "name": name}


def make_rmat(log2_n: int, avg_deg: float, seed: int = 0, device="cpu", values="gcn",
              abcd=(0.57, 0.19, 0.19, 0.05)):
    """R-MAT edge list -> CSR (row = destination), duplicates kept (multigraph)."""
    n = 1 << log2_n
    e = int(n * avg_deg)
    assert e < 2**31
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    a, b, c, _ = abcd
    rows = torch.zeros(e, dtype=torch.int64, device=device)
    cols = torch.zeros(e, dtype=torch.int64, device=device)
    for _ in range(log2_n):
        u = torch.rand(e, generator=gen, device=device)
        right = ((u >= a) & (u < a + b)) | (u >= a + b + c)
        down = u >= a + b
        rows = (rows << 1) | down.long()
        cols = (cols << 1) | right.long()
    key, _ = torch.sort(rows * n + cols)
    rows, cols = key // n, key % n
    deg = torch.bincount(rows, minlength=n)
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    torch.cumsum(deg, 0, out=indptr[1:])
    return finish_graph(indptr.to(torch.int32), cols.to(torch.int32), n, gen, values, f"rmat{log2_n}")


def algorithmic_bytes(n_rows: int, e: int, w: int, k: int, d: int = 256, n_cbsr=None):
    """BASELINE.md section 3 / SURVEY 8(d): the byte counts roofline.achieved is computed from."""
    n_cbsr = n_rows if n_cbsr is None else n_cbsr
    fwd = e * (8 + 5 * k) + 16 * w + 4 * n_rows * d
    bwd = e * (8 + 5 * k) + 16 * w + 4 * n_rows * d + 4 * n_cbsr * k
    return {"fwd": fwd, "bwd": bwd, "topk": 4 * n_cbsr * d + 5 * n_cbsr * k,
            "meta": 4 * (n_rows + 1) + 16 * w,
            "compulsory_fwd": e * 8 + 16 * w + 5 * n_cbsr * k + 4 * n_rows * d}
