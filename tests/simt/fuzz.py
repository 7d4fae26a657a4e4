#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: random-parameter fuzzing of the kernel sources under the SIMT interpreter
against the oracle (forward, backward, TopK, work list, layout kernels).

    python tests/simt/fuzz.py [--seconds 120] [--seed 0]

Prints one line per failure with the parameters that reproduce it; exit code 1 if any.
"""
import argparse
import os
import sys
import time
import traceback

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "spgemm-new_b200")]
import simt_lib as S  # noqa: E402
from oracle import oracle as O  # noqa: E402

RTOL, ATOL = 1e-5, 1e-6


def one(rng, it):
    d = int(rng.choice([4, 8, 12, 32, 36, 64, 100, 128, 200, 256]))
    k = int(rng.choice([1, 2, 3, 4, 8, 16, 18, 20, 32, 33, 64, 96, 128, 160, 256]))
    k = min(k, d)
    n = int(rng.choice([1, 2, 31, 33, 100, 257]))
    n_cols = n if rng.random() < 0.7 else int(rng.integers(1, 400))
    avg = float(rng.choice([0.5, 3.0, 20.0, 70.0, 300.0]))
    nz = int(rng.choice([64, 64, 64, 1, 5, 63, 65, 100, 200]))
    shape = int(rng.choice([0, 0, 1, 2, 3, 8, 12]))
    copies = int(rng.choice([0, 0, 1, 2, 4, 8, 16]))
    tpw_f, tpw_b = int(rng.choice([-1, 1, 2, 3, 16, 100])), int(rng.choice([-1, 1, 2, 16]))
    extra_rows = int(rng.choice([0, 0, 1, 9]))
    accumulate = bool(rng.random() < 0.2)
    al_d, al_i, al_x = int(rng.choice([16, 16, 4])), int(rng.choice([16, 4, 1])), int(rng.choice([16, 16, 4]))
    selector = str(rng.choice(["topk", "random"]))
    eager = bool(rng.random() < 0.5)
    params = dict(it=it, d=d, k=k, n=n, n_cols=n_cols, avg=avg, nz=nz, shape=shape, copies=copies, tpw_f=tpw_f,
                  tpw_b=tpw_b, extra_rows=extra_rows, accumulate=accumulate, al_d=al_d, al_i=al_i, al_x=al_x,
                  selector=selector, eager=eager)
    seed = int(rng.integers(1 << 30))
    r = np.random.default_rng(seed)
    indptr, idx = O.synth_graph(n, avg, seed=seed, empty_frac=float(rng.choice([0.0, 0.3])), n_cols=n_cols)
    e = idx.shape[0]
    # forward/backward operands are positive like the GPU parity tests' (the fp32 rtol 1e-5 / atol 1e-6
    # bar is stated for sums without cancellation); TopK sees signed data with ties
    val = r.random(e, dtype=np.float32)
    x = r.standard_normal((n_cols, d)).astype(np.float32)
    if r.random() < 0.3:
        x = np.round(x)  # ties
    dy = r.random((n, d), dtype=np.float32)
    try:
        w4 = O.worklist(indptr, nz)
        got_w4 = S.worklist(indptr, nz)
        assert np.array_equal(got_w4, w4.reshape(-1)), "worklist"
        ov, oi = O.topk_cbsr(x, k)
        v, i = S.topk_f32(x, k)
        assert np.array_equal(i, oi) and np.array_equal(v.view(np.uint32), ov.view(np.uint32)), "topk"
        # fc_self on the CBSR operand (cbsr_dense.cu) against the dense Linear on the scattered twin
        c = int(rng.choice([4, 8, 48, 128, 132, 256, 260]))
        params["c"] = c
        w = (r.standard_normal((c, d)) * 0.2).astype(np.float32)
        bias = r.standard_normal(c).astype(np.float32) if r.random() < 0.5 else None
        m = np.ascontiguousarray(w.T)
        terms = np.abs(O.cbsr_scatter(ov, oi, d)).astype(np.float64) @ np.abs(m).astype(np.float64) + (0 if bias is None else np.abs(bias))
        ylin = S.cbsr_matmul(ov, oi, m, bias=bias)
        wl = O.cbsr_linear(ov, oi, w, bias)
        assert (np.abs(ylin - wl) <= np.maximum(ATOL + RTOL * np.abs(wl), (k + 3) * 2.0 ** -24 * terms)).all(), "cbsr_matmul"
        gl = r.standard_normal((n_cols, c)).astype(np.float32)
        wg = O.cbsr_linear_grad_values(gl, oi, w)
        gterms = np.take_along_axis(np.abs(gl).astype(np.float64) @ np.abs(w).astype(np.float64), oi.astype(np.int64), axis=1)
        dvl = S.cbsr_sampled_dot(gl, m, oi)
        assert (np.abs(dvl - wg) <= np.maximum(ATOL + RTOL * np.abs(wg), (c + 2) * 2.0 ** -24 * gterms)).all(), "cbsr_sampled_dot"
        ov = np.abs(ov)
        if selector == "random":
            oi = np.argsort(r.random((n_cols, d)), axis=1)[:, :k].astype(np.uint8)
            ov = r.random((n_cols, k), dtype=np.float32)
        S.set_forward_depth(shape); S.set_forward_copies(copies); S.set_tuning(tpw_f, tpw_b); S.set_cp_async_eager(eager)
        n_out = n + extra_rows
        want = O.spgemm_forward(w4, idx, val, ov, oi, d, n_out=n_out)
        init = r.random((n_out, d), dtype=np.float32) if accumulate else None
        y = S.spgemm_forward(w4, idx, val, ov, oi, n_out, d=d, accumulate=accumulate, out_init=init,
                             data_align=al_d, index_align=al_i)
        ref = want + (init if accumulate else 0)
        err = np.abs(y.astype(np.float64) - ref) - RTOL * np.abs(ref)
        assert np.isfinite(y).all() and err.max(initial=0.0) <= ATOL, f"forward: worst excess {err.max():.3g} at {np.unravel_index(err.argmax(), err.shape)}"
        wantg = O.sspmm_backward(w4, idx, val, dy, oi, n_cbsr=n_cols)
        # a dXc element is an fp32 sum of (edges into its node) terms added in arbitrary order: the
        # rtol 1e-5 bar is stated for the graph shapes of BASELINE.json (hundreds of terms); widen it
        # like sqrt(#terms) for the hub columns this fuzzer can draw (e.g. 77 k edges into 3 nodes)
        fan_in = max(1.0, e / max(n_cols, 1))
        RT_B = RTOL * max(1.0, (fan_in / 1000.0) ** 0.5 * 4.0)
        g = S.sspmm_backward(w4, idx, val, dy, oi, n_cbsr=n_cols, dxc_align=al_x, index_align=al_i)
        err = np.abs(g.astype(np.float64) - wantg) - RT_B * np.abs(wantg)
        assert np.isfinite(g).all() and err.max(initial=0.0) <= ATOL, f"backward: worst excess {err.max():.3g} at {np.unravel_index(err.argmax(), err.shape)}"
        # the bank plan: planned forward where it is built (k = 32, 64), the backward decoding its columns for any k
        sa = S.cbsr_bank_plan(oi)
        if k in (32, 64):
            S.set_forward_depth(int(rng.choice([0, 1, 2])))
            yp = S.spgemm_forward_planned(w4, idx, val, ov, sa, n_out, d=d, accumulate=accumulate, out_init=init,
                                          data_align=al_d if al_d >= 4 else 4, addr_align=int(rng.choice([16, 2])))
            err = np.abs(yp.astype(np.float64) - ref) - RTOL * np.abs(ref)
            assert np.isfinite(yp).all() and err.max(initial=0.0) <= ATOL, f"planned forward: worst excess {err.max():.3g}"
        gp = S.sspmm_backward_planned(w4, idx, val, dy, sa, n_cbsr=n_cols, dxc_align=al_x,
                                      addr_align=int(rng.choice([16, 8, 2])))
        err = np.abs(gp.astype(np.float64) - wantg) - RT_B * np.abs(wantg)
        assert np.isfinite(gp).all() and err.max(initial=0.0) <= ATOL, f"planned backward: worst excess {err.max():.3g}"
        dense = S.cbsr_scatter(ov, oi, d)
        assert np.array_equal(dense, O.cbsr_scatter(ov, oi, d)), "scatter"
        assert np.array_equal(S.cbsr_gather(dense, oi), ov), "gather"
        return None
    except Exception as ex:  # noqa: BLE001
        return params, "".join(traceback.format_exception_only(type(ex), ex)).strip().splitlines()[-1][:300]
    finally:
        S.set_forward_depth(0); S.set_forward_copies(0); S.set_tuning(-1, -1); S.set_cp_async_eager(False)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    t0, it, bad = time.time(), 0, 0
    while time.time() - t0 < a.seconds:
        res = one(rng, it)
        it += 1
        if res:
            bad += 1
            print("FAIL", res[0], "::", res[1], flush=True)
            if bad >= 20:
                break
    print(f"{it} cases, {bad} failures, {time.time() - t0:.0f} s")
    sys.exit(1 if bad else 0)
