#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: how sharp are the SIMT tests?  Each mutant below is a one-line change to
a COPY of spgemm-new_b200/csrc (a dropped barrier, an off-by-one, a wrong mask ...); the copy is
built for the interpreter and tests/test_simt_kernels.py must FAIL on it.  Prints one line per
mutant; exit code 1 if a mutant survives.

    python tests/simt/mutation_check.py [name ...]
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "spgemm-new_b200", "csrc")

# name: (file, old, new, occurrence)
MUTANTS = {
    "fwd-stage-no-barrier-after-write": ("spgemm_forward.cu", "    stage[lane + 32] = make_int2(r.i1, __float_as_int(r.v1));\n    __syncwarp();",
                                         "    stage[lane + 32] = make_int2(r.i1, __float_as_int(r.v1));", 0),
    "fwd-batch-runs-past-the-task": ("spgemm_forward.cu", "for (; j + STEP <= len; j += STEP) {  // every lane has an edge",
                                     "for (; j + STEP <= len + 1; j += STEP) {", 0),
    "fwd-shared-row-as-plain-store": ("spgemm_forward.cu", "const bool first_shared = accumulate || (prev_row == first_row);", "const bool first_shared = accumulate;", 0),
    "fwd-gap-rows-not-zeroed": ("spgemm_forward.cu", "if (!accumulate) fwd_zero_rows(out, (int64_t)cur_row + 1, info.x, d, lane);", "", 0),
    "ring-no-barrier-after-wait": ("spgemm_forward.cu", "        __syncwarp();            // ... and so have the other lanes'\n", "", 0),
    "ring-waits-for-too-little": ("spgemm_forward.cu", "            cp_async_wait<1>();  // this lane's chunks of group g have landed ...", "            cp_async_wait<2>();", 0),
    "ring-selector-chunk-offset": ("spgemm_forward.cu", "sp_index + (int64_t)st[e].x * K + 16 * cs);", "sp_index + (int64_t)st[e].x * K + 16 * cs + 16);", 0),
    "bwd-stage-no-barrier": ("sspmm_backward.cu", "    stage[lane + 32] = make_int2(r.i1, __float_as_int(r.v1));\n    __syncwarp();", "    stage[lane + 32] = make_int2(r.i1, __float_as_int(r.v1));", 0),
    "bwd-dy-restage-no-barrier": ("sspmm_backward.cu", "            cur_row = info.x;\n            __syncwarp();", "            cur_row = info.x;", 0),
    "bwd-selector-of-wrong-node": ("sspmm_backward.cu", "reinterpret_cast<const unsigned *>(sp_index + (int64_t)ev[u].x * k) + q[u]);", "reinterpret_cast<const unsigned *>(sp_index + (int64_t)ev[u].x * k) + (q[u] ^ 1));", 0),
    "topk-emit-no-barrier": ("topk_select.cuh", "    __syncwarp();\n    unsigned char *oi = out_idx + row * k + lane;", "    unsigned char *oi = out_idx + row * k + lane;", 0),
    "topk-tie-order": ("topk_select.cuh", "if (r0 < need) sel |= (1u << i);", "if (r0 <= need) sel |= (1u << i);", 0),
    # the fast path's proof: only a probe that separates EXACTLY k elements may end the search
    "topk-proof-accepts-k-plus-1": ("topk_cbsr.cu", "            if (g == 0) {\n                done = true;", "            if (g == 0 || g == 1) {\n                done = true;", 0),
    # level snapping of the robust search: the element below a probe IS the k-th largest once its count reaches k
    "topk-snap-up-skips-an-element": ("topk_select.cuh", "                    L = m + 1u;\n                    cL = c1;", "                    L = m + 2u;\n                    cL = c1;", 0),
    "topk-row-bound": ("topk_cbsr.cu", "if (row >= n) return;  // warp-uniform; no CTA barrier below\n\n", "if (row > n) return;\n\n", 0),
    # bank plan and planned kernels
    "plan-home-word-formula": ("common.cuh", "return h == 0 ? c : 256u + (c & 7u) * 32u + (c >> 3);", "return h == 0 ? c : 256u + (c & 7u) * 32u + ((c >> 3) ^ 1u);", 0),
    "plan-word-column-decode": ("common.cuh", "return w < 256u ? w : (((w & 31u) << 3) | ((w >> 5) & 7u));", "return w < 256u ? w : (((w & 31u) << 3) | ((w >> 5) & 3u));", 0),
    "plan-moved-holder-keeps-its-home": ("cbsr_plan.cu", "            home ^= 1;\n            move[lane] = 0;", "            move[lane] = 0;", 0),
    "plan-no-barrier-before-reading-the-owner-table": ("cbsr_plan.cu", "    if (valid) owner[b0] = (unsigned char)lane;\n    __syncwarp();", "    if (valid) owner[b0] = (unsigned char)lane;", 0),
    "planned-flush-skips-a-home": ("spgemm_forward.cu", "        for (int h = 1; h < H; ++h) {\n            const float *p = acc + h * 256;", "        for (int h = 1; h < H - 1; ++h) {\n            const float *p = acc + h * 256;", 0),
    "planned-flush-clears-before-all-lanes-read": ("spgemm_forward.cu", "    __syncwarp();  // every lane has read all homes before anybody clears a word of them\n", "", 0),
    # NVLink peer path (tests/test_simt_python_layer.py -k peer): the stand-in for the TMA unit lands data late
    "peer-tma-stage-reused-before-its-stores-read-it": ("peer.cuh", 'if (chunks_issued >= PEER_TMA_STAGES) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");', "", "peer"),
    "peer-tma-store-before-the-load-landed": ("peer.cuh", "        mbar_wait(mbar0 + 8 * st, (phase >> st) & 1u);\n", "", "peer"),
    "peer-flag-raised-for-the-wrong-rank": ("peer.cuh", "st_release_sys(a.flags[p] + PEER_FLAG_ARRIVED + a.rank, a.epoch);", "st_release_sys(a.flags[p] + PEER_FLAG_ARRIVED + p, a.epoch);", "peer"),
    "peer-halo-rows-of-the-wrong-peer": ("peer.cuh", "    const int *rows = a.send_rows[p];\n    const long long n = a.send_count[p];", "    const int *rows = a.send_rows[p];\n    const long long n = a.send_count[p] - 1;", "peer"),
    # CBSR x dense products (cbsr_dense.cu)
    "cbsr-dense-no-barrier-after-staging": ("cbsr_dense.cu", "            __syncwarp();\n            float4 acc = b4;", "            float4 acc = b4;", 0),
    "cbsr-dense-no-barrier-before-restaging": ("cbsr_dense.cu", "            __syncwarp();  // the staging area is rewritten for the next row\n", "", 0),
    "cbsr-dense-tile-reload-no-barrier": ("cbsr_dense.cu", "        __syncthreads();  // everyone is done with the previous tile\n", "", 0),
    "cbsr-dense-transpose-keeps-the-sent-half": ("cbsr_dense.cu", "const float keep = upper ? p[i + half] : p[i];", "const float keep = upper ? p[i] : p[i + half];", 0),
    "cbsr-dense-later-tiles-overwrite": ("cbsr_dense.cu", "const bool first = (col0 == 0) && !accumulate;", "const bool first = !accumulate;", 0),
    "worklist-scan-no-barrier": ("worklist.cu", "    if (lane == 31) warp_sums[warp] = inc;\n    __syncthreads();", "    if (lane == 31) warp_sums[warp] = inc;", 0),
    "worklist-last-chunk-length": ("worklist.cu", "rec.z = min(nz, deg - nz * c_in_row);", "rec.z = nz;", 0),
}

# Mutants that SURVIVE because the change does not alter behaviour (listed so nobody re-adds them):
# with MAXK_FWD_SYNCWARP=1 every fwd_accumulate ends in a __syncwarp, which already orders the
# previous batch's staging reads before the next fwd_stage_edges write and the last accumulate
# before fwd_flush_row -- the leading barriers of those two functions are redundant (they stay:
# they cost nothing measurable and keep the functions correct on their own).
EQUIVALENT = {
    # level snapping of the robust TopK search, element below a probe with count == k: not ending there leaves
    # H = m with cH = k; every later end of the search (a probe with count k, or H - L == 1 with "ties at L" whose
    # strictly greater elements are exactly those k) selects the same set -- slower, not wrong
    "topk-snap-down-keeps-the-level": ("topk_select.cuh", "                    if (cm >= k) {  // the elements above m", "                    if (cm > k) {  // the elements above m", 0),
    # every row starts with warp-wide REDUX collectives (row minimum / maximum): all lanes have left the previous
    # row's emission reads before any lane can reach the next row's staging writes; the barrier stays as
    # documentation of the reuse
    "topk-rows-no-barrier-between-rows": ("topk_cbsr.cu", "        __syncwarp();  // the staging area is reused by the next row\n", "", 0),
    "fwd-stage-no-barrier-before-write": ("spgemm_forward.cu", "    __syncwarp();  // everyone is done reading the previous batch\n", "", 0),
    "fwd-flush-no-leading-barrier": ("spgemm_forward.cu", "    constexpr int R = C::R;\n    __syncwarp();", "    constexpr int R = C::R;", 0),
    # ... as is the barrier between the ring accumulate and the re-issue into the same half
    "ring-no-barrier-before-reissue": ("spgemm_forward.cu", "        __syncwarp();            // everyone is done reading this half\n", "", 0),
    # ... and the trailing barrier of fwd_flush_row is followed by fwd_stage_edges' leading one
    "fwd-flush-no-trailing-barrier": ("spgemm_forward.cu", "        }\n    }\n    __syncwarp();\n}\n\n// coalesced fetch", "        }\n    }\n}\n\n// coalesced fetch", 0),
}


def run(name, spec):
    fname, old, new, where = spec
    target = [os.path.join(ROOT, "tests", "test_simt_kernels.py")]
    if where == "peer":
        target = [os.path.join(ROOT, "tests", "test_simt_python_layer.py"), "-k", "peer"]
    tmp = tempfile.mkdtemp(prefix="simt_mut_")
    try:
        src = os.path.join(tmp, "csrc")
        shutil.copytree(CSRC, src)
        path = os.path.join(src, fname)
        text = open(path).read()
        if text.count(old) < 1:
            return "PATTERN NOT FOUND (update the mutant)"
        open(path, "w").write(text.replace(old, new, 1))
        env = dict(os.environ, SIMT_CSRC=src, SIMT_OUT=os.path.join(tmp, "out"))
        try:
            r = subprocess.run([sys.executable, "-m", "pytest", *target, "-x", "-q", "-p", "no:cacheprovider"],
                               env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
        except subprocess.TimeoutExpired:
            # ranks of the peer path that wait for a flag nobody raises spin until the host-set limit (minutes)
            return "killed  (no result within 900 s: the mutant hangs the ranks)"
        if r.returncode == 0:
            return "SURVIVED"
        last = [ln for ln in r.stdout.splitlines() if ln.startswith(("FAILED", "ERROR"))]
        return "killed  (" + (last[0][:110] if last else "rc=%d" % r.returncode) + ")"
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    names = sys.argv[1:] or [n for n, s in MUTANTS.items() if s]
    bad = 0
    for n in names:
        res = run(n, MUTANTS.get(n) or EQUIVALENT[n])
        bad += not res.startswith("killed")
        print(f"{n:40s} {res}", flush=True)
    sys.exit(1 if bad else 0)
