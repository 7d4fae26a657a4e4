"""TEST INFRASTRUCTURE: numpy front end of tests/simt/_build/libmaxk_simt.so -- the CUDA kernel
sources of spgemm-new_b200/csrc compiled for the host and run under the SIMT test interpreter
(tests/simt/).  It exists so that the `-m "not gpu"` suite can check the kernels' logic and
memory safety against the oracle in a container without a GPU.  The product never loads this
library: `maxk_b200._cabi` binds lib/libmaxk_b200.so only and raises if that is missing.

Every array handed to a kernel lives in its own mapping between two PROT_NONE guard pages, with
its END flush against the upper guard (up to the alignment the ABI asks for) and canary bytes in
the slack, so a read or write outside the array either faults (reported by the interpreter with
the address and the CUDA thread) or trips the canary check after the call.
"""
import ctypes
import importlib.util
import mmap
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PAGE = mmap.PAGESIZE
_libc = ctypes.CDLL(None, use_errno=True)
_libc.mprotect.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
_lib = None


class SimtError(AssertionError):
    pass


def load():
    global _lib
    if _lib is None:
        spec = importlib.util.spec_from_file_location("simt_build", os.path.join(HERE, "simt", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        L = ctypes.CDLL(mod.build())
        L.maxk_last_error.restype = ctypes.c_char_p
        L.simt_last_failure.restype = ctypes.c_char_p
        L.maxk_launch_count.restype = ctypes.c_int64
        L.maxk_worklist_workspace_bytes.restype = ctypes.c_int64
        L.maxk_worklist_workspace_bytes.argtypes = [ctypes.c_int64]
        _lib = L
    return _lib


class Guarded:
    """An array between guard pages.  `data`: array to copy in, or None with shape/dtype for an
    output (poisoned: NaN for floats, 0xEE bytes otherwise)."""

    def __init__(self, data=None, shape=None, dtype=None, align=16, name="buffer"):
        if data is not None:
            data = np.ascontiguousarray(data)
            shape, dtype = data.shape, data.dtype
        dtype = np.dtype(dtype)
        self.name, self.nbytes = name, int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        body = -(-(self.nbytes + align) // PAGE) * PAGE or PAGE
        self._mm = mmap.mmap(-1, body + 2 * PAGE)
        self._base = ctypes.addressof(ctypes.c_char.from_buffer(self._mm))
        lo, hi = self._base + PAGE, self._base + PAGE + body
        self.addr = (hi - self.nbytes) & ~(align - 1)
        self._lo, self._hi = lo, hi
        raw = np.frombuffer(self._mm, dtype=np.uint8, count=body, offset=PAGE)
        raw[:] = 0xA5
        self._raw = raw
        off = self.addr - lo
        self.array = raw[off:off + self.nbytes].view(dtype).reshape(shape)
        if data is not None:
            self.array[...] = data
        elif dtype.kind == "f":
            self.array[...] = np.nan
        else:
            raw[off:off + self.nbytes] = 0xEE
        for page in (self._base, hi):
            if _libc.mprotect(page, PAGE, 0) != 0:
                raise OSError(ctypes.get_errno(), "mprotect")

    @property
    def ptr(self):
        return ctypes.c_void_p(self.addr) if self.nbytes or True else None

    def check(self):
        off = self.addr - self._lo
        if not (self._raw[:off] == 0xA5).all():
            raise SimtError(f"{self.name}: bytes BEFORE the array were written")
        if not (self._raw[off + self.nbytes:] == 0xA5).all():
            raise SimtError(f"{self.name}: bytes PAST the end of the array were written")

    def describe(self, address):
        if self._base <= address < self._hi + PAGE:
            return f"{self.name} [{self.addr:#x}, {self.addr + self.nbytes:#x}) offset {address - self.addr:+d}"
        return None

    def close(self):
        for page in (self._base, self._hi):
            _libc.mprotect(page, PAGE, 3)
        self.array = self._raw = None
        try:
            self._mm.close()
        except BufferError:
            pass


def _call(fn, bufs, *args):
    """call an ABI function; raise with the library's (and the interpreter's) message on failure"""
    L = load()
    rc = fn(*args)
    if rc != 0:
        msg = L.maxk_last_error().decode()
        fail = L.simt_last_failure().decode()
        where = ""
        if "address 0x" in fail:
            a = int(fail.rsplit("address ", 1)[1].split()[0], 16)
            where = "; ".join(filter(None, (b.describe(a) for b in bufs)))
        raise SimtError(f"rc={rc}: {msg} {('[' + where + ']') if where else ''}")
    for b in bufs:
        b.check()


def _run(fn, ins, outs, build_args):
    """ins/outs: dict name -> Guarded.  build_args(ptrs) -> argument tuple."""
    bufs = list(ins.values()) + list(outs.values())
    try:
        _call(fn, bufs, *build_args({n: b.ptr for n, b in {**ins, **outs}.items()}))
        return {n: b.array.copy() for n, b in outs.items()}
    finally:
        for b in bufs:
            b.close()


i32, i64 = ctypes.c_int32, ctypes.c_int64


def set_tuning(fwd=0, bwd=0):
    load().maxk_set_tuning(int(fwd), int(bwd))


def set_forward_depth(depth=0):
    load().maxk_set_forward_depth(int(depth))


def set_topk_group(lanes=-1):
    rc = load().maxk_set_topk_group(int(lanes))
    assert rc == 0, rc


def set_forward_copies(copies=0):
    load().maxk_set_forward_copies(int(copies))


def set_cp_async_eager(eager):
    """cp.async data lands immediately (True) or as late as PTX allows, at the issuer's wait (False, default)."""
    load().simt_set_cp_async_eager(1 if eager else 0)


def launch_count():
    return load().maxk_launch_count()


def worklist(indptr, warp_max_nz=64):
    L = load()
    indptr = np.ascontiguousarray(indptr, np.int32)
    n = indptr.shape[0] - 1
    ws_bytes = max(int(L.maxk_worklist_workspace_bytes(n)), 16)
    ip = Guarded(indptr, name="indptr", align=4)
    ws = Guarded(shape=(ws_bytes,), dtype=np.uint8, name="workspace")
    count = ctypes.c_int64(-1)
    try:
        _call(L.maxk_worklist_count, [ip, ws], ip.ptr, i64(n), i32(warp_max_nz), ws.ptr, ctypes.byref(count), None)
        w = int(count.value)
        out = Guarded(shape=(max(w, 0) * 4,), dtype=np.int32, name="warp4")
        try:
            _call(L.maxk_worklist_fill, [ip, ws, out], ip.ptr, i64(n), i32(warp_max_nz), ws.ptr, out.ptr, i64(w), None)
            return out.array.copy()
        finally:
            out.close()
    finally:
        ip.close()
        ws.close()


def topk_f32(x, k, want_dense=False, want_cbsr=True, out_align=(4, 1)):
    L = load()
    x = np.ascontiguousarray(x, np.float32)
    n, d = x.shape
    ins = {"x": Guarded(x, name="x")}
    outs = {}
    if want_cbsr:
        outs["sp_data"] = Guarded(shape=(n, k), dtype=np.float32, name="sp_data", align=out_align[0])
        outs["sp_index"] = Guarded(shape=(n, k), dtype=np.uint8, name="sp_index", align=out_align[1])
    if want_dense:
        outs["dense"] = Guarded(shape=(n, d), dtype=np.float32, name="dense")
    r = _run(L.maxk_topk_cbsr_f32, ins, outs,
             lambda p: (p["x"], i64(n), i32(d), i32(k), p.get("sp_data"), p.get("sp_index"), p.get("dense"), None))
    return tuple(r[n_] for n_ in ("sp_data", "sp_index", "dense") if n_ in r)


def topk_u8(x, k):
    L = load()
    x = np.ascontiguousarray(x, np.uint8)
    n, d = x.shape
    r = _run(L.maxk_topk_cbsr_u8, {"x": Guarded(x, name="x", align=4)},
             {"value": Guarded(shape=(n, k), dtype=np.uint8, name="value", align=1),
              "index": Guarded(shape=(n, k), dtype=np.uint8, name="index", align=1)},
             lambda p: (p["x"], i64(n), i32(d), i32(k), p["value"], p["index"], None))
    return r["value"], r["index"]


def spgemm_forward(w4, idx, val, sp_data, sp_index, n_out, d=256, accumulate=False, out_init=None,
                   data_align=16, index_align=4):
    L = load()
    w4 = np.ascontiguousarray(w4, np.int32).reshape(-1)
    k = sp_data.shape[1]
    ins = {"warp4": Guarded(w4, name="warp4"), "idx": Guarded(np.asarray(idx, np.int32), name="indices", align=4),
           "val": Guarded(np.asarray(val, np.float32), name="values", align=4),
           "sp_data": Guarded(np.asarray(sp_data, np.float32), name="sp_data", align=data_align),
           "sp_index": Guarded(np.asarray(sp_index, np.uint8), name="sp_index", align=index_align)}
    out = Guarded(out_init, name="out") if out_init is not None else Guarded(shape=(n_out, d), dtype=np.float32, name="out")
    r = _run(L.maxk_spgemm_forward_acc, ins, {"out": out},
             lambda p: (p["warp4"], i64(w4.shape[0] // 4), p["idx"], p["val"], p["sp_data"], p["sp_index"], i32(k),
                        p["out"], i64(n_out), i32(d), i32(1 if accumulate else 0), None))
    return r["out"]


def sspmm_backward(w4, idx, val, dy, sp_index, n_cbsr=None, accumulate=False, out_init=None, dxc_align=16,
                   index_align=4):
    L = load()
    w4 = np.ascontiguousarray(w4, np.int32).reshape(-1)
    sp_index = np.asarray(sp_index, np.uint8)
    n_cbsr = sp_index.shape[0] if n_cbsr is None else n_cbsr
    k, d = sp_index.shape[1], dy.shape[1]
    ins = {"warp4": Guarded(w4, name="warp4"), "idx": Guarded(np.asarray(idx, np.int32), name="indices", align=4),
           "val": Guarded(np.asarray(val, np.float32), name="values", align=4),
           "dy": Guarded(np.asarray(dy, np.float32), name="dy"),
           "sp_index": Guarded(sp_index, name="sp_index", align=index_align)}
    out = (Guarded(out_init, name="dxc", align=dxc_align) if out_init is not None
           else Guarded(shape=(n_cbsr, k), dtype=np.float32, name="dxc", align=dxc_align))
    r = _run(L.maxk_sspmm_backward, ins, {"dxc": out},
             lambda p: (p["warp4"], i64(w4.shape[0] // 4), p["idx"], p["val"], p["dy"], p["sp_index"], i32(k), i32(d),
                        p["dxc"], i64(n_cbsr), i32(1 if accumulate else 0), None))
    return r["dxc"]


def cbsr_scatter(sp_data, sp_index, d):
    L = load()
    n, k = sp_index.shape
    r = _run(L.maxk_cbsr_scatter, {"sp_data": Guarded(np.asarray(sp_data, np.float32), name="sp_data", align=4),
                                   "sp_index": Guarded(np.asarray(sp_index, np.uint8), name="sp_index", align=1)},
             {"dense": Guarded(shape=(n, d), dtype=np.float32, name="dense")},
             lambda p: (p["sp_data"], p["sp_index"], i64(n), i32(k), i32(d), p["dense"], None))
    return r["dense"]


def cbsr_gather(dense, sp_index):
    L = load()
    n, k = sp_index.shape
    d = dense.shape[1]
    r = _run(L.maxk_cbsr_gather, {"dense": Guarded(np.asarray(dense, np.float32), name="dense"),
                                  "sp_index": Guarded(np.asarray(sp_index, np.uint8), name="sp_index", align=1)},
             {"sp_data": Guarded(shape=(n, k), dtype=np.float32, name="sp_data", align=4)},
             lambda p: (p["dense"], p["sp_index"], i64(n), i32(k), i32(d), p["sp_data"], None))
    return r["sp_data"]


def cbsr_matmul(sp_data, sp_index, m, bias=None, accumulate=False, out_init=None):
    L = load()
    n, k = sp_index.shape
    d, c = m.shape
    ins = {"sp_data": Guarded(np.asarray(sp_data, np.float32), name="sp_data", align=4),
           "sp_index": Guarded(np.asarray(sp_index, np.uint8), name="sp_index", align=1),
           "m": Guarded(np.asarray(m, np.float32), name="m")}
    if bias is not None:
        ins["bias"] = Guarded(np.asarray(bias, np.float32), name="bias")
    out = Guarded(np.asarray(out_init, np.float32), name="out") if out_init is not None else \
        Guarded(shape=(n, c), dtype=np.float32, name="out")
    r = _run(L.maxk_cbsr_matmul, ins, {"out": out},
             lambda p: (p["sp_data"], p["sp_index"], i64(n), i32(k), p["m"], i32(d), i32(c), p.get("bias"), p["out"],
                        i32(1 if accumulate else 0), None))
    return r["out"]


def cbsr_axpy(y, sp_data, sp_index, alpha=1.0, alpha_dev=None):
    L = load()
    n, k = sp_index.shape
    ins = {"sp_data": Guarded(np.asarray(sp_data, np.float32), name="sp_data", align=4),
           "sp_index": Guarded(np.asarray(sp_index, np.uint8), name="sp_index", align=1)}
    if alpha_dev is not None:
        ins["alpha"] = Guarded(np.asarray([alpha_dev], np.float32), name="alpha", align=4)
    r = _run(L.maxk_cbsr_axpy, ins, {"y": Guarded(np.asarray(y, np.float32), name="y", align=4)},
             lambda p: (p["sp_data"], p["sp_index"], i64(n), i32(k), i32(y.shape[1]), p.get("alpha"),
                        ctypes.c_float(alpha), p["y"], None))
    return r["y"]


def cbsr_sampled_dot(g, m, sp_index, accumulate=False, out_init=None):
    L = load()
    n, k = sp_index.shape
    d, c = m.shape
    out = Guarded(np.asarray(out_init, np.float32), name="out", align=4) if out_init is not None else \
        Guarded(shape=(n, k), dtype=np.float32, name="out", align=4)
    r = _run(L.maxk_cbsr_sampled_dot, {"g": Guarded(np.asarray(g, np.float32), name="g"),
                                       "m": Guarded(np.asarray(m, np.float32), name="m"),
                                       "sp_index": Guarded(np.asarray(sp_index, np.uint8), name="sp_index", align=1)},
             {"out": out},
             lambda p: (p["g"], i64(n), i32(c), p["m"], i32(d), p["sp_index"], i32(k), p["out"],
                        i32(1 if accumulate else 0), None))
    return r["out"]


def cbsr_mask(dense, sp_index):
    L = load()
    n, k = sp_index.shape
    d = dense.shape[1]
    r = _run(L.maxk_cbsr_mask, {"in": Guarded(np.asarray(dense, np.float32), name="in"),
                                "sp_index": Guarded(np.asarray(sp_index, np.uint8), name="sp_index", align=1)},
             {"out": Guarded(shape=(n, d), dtype=np.float32, name="out")},
             lambda p: (p["in"], p["sp_index"], i64(n), i32(k), i32(d), p["out"], None))
    return r["out"]


def cbsr_pack(sp_data, sp_index):
    L = load()
    n, k = sp_index.shape
    rb = int(L.maxk_cbsr_record_bytes(i32(k)))
    r = _run(L.maxk_cbsr_pack, {"sp_data": Guarded(np.asarray(sp_data, np.float32), name="sp_data"),
                                "sp_index": Guarded(np.asarray(sp_index, np.uint8), name="sp_index", align=4)},
             {"records": Guarded(shape=(n, rb), dtype=np.uint8, name="records")},
             lambda p: (p["sp_data"], p["sp_index"], i64(n), i32(k), p["records"], None))
    return r["records"]


def cbsr_unpack(records, n, k):
    L = load()
    r = _run(L.maxk_cbsr_unpack, {"records": Guarded(np.asarray(records, np.uint8), name="records")},
             {"sp_data": Guarded(shape=(n, k), dtype=np.float32, name="sp_data"),
              "sp_index": Guarded(shape=(n, k), dtype=np.uint8, name="sp_index", align=4)},
             lambda p: (p["records"], i64(n), i32(k), p["sp_data"], p["sp_index"], None))
    return r["sp_data"], r["sp_index"]


def cbsr_bank_plan(sp_index):
    L = load()
    sp_index = np.ascontiguousarray(sp_index, np.uint8)
    n, k = sp_index.shape
    r = _run(L.maxk_cbsr_bank_plan, {"sp_index": Guarded(sp_index, name="sp_index", align=1)},
             {"sp_addr": Guarded(shape=(n, k), dtype=np.uint16, name="sp_addr", align=2)},
             lambda p: (p["sp_index"], i64(n), i32(k), p["sp_addr"], None))
    return r["sp_addr"]


def spgemm_forward_planned(w4, idx, val, sp_data, sp_addr, n_out, d=256, accumulate=False, out_init=None,
                           data_align=16, addr_align=16):
    L = load()
    w4 = np.ascontiguousarray(w4, np.int32).reshape(-1)
    k = sp_data.shape[1]
    ins = {"warp4": Guarded(w4, name="warp4"), "idx": Guarded(np.asarray(idx, np.int32), name="indices", align=4),
           "val": Guarded(np.asarray(val, np.float32), name="values", align=4),
           "sp_data": Guarded(np.asarray(sp_data, np.float32), name="sp_data", align=data_align),
           "sp_addr": Guarded(np.asarray(sp_addr, np.uint16), name="sp_addr", align=addr_align)}
    out = Guarded(out_init, name="out") if out_init is not None else Guarded(shape=(n_out, d), dtype=np.float32, name="out")
    r = _run(L.maxk_spgemm_forward_planned, ins, {"out": out},
             lambda p: (p["warp4"], i64(w4.shape[0] // 4), p["idx"], p["val"], p["sp_data"], p["sp_addr"], i32(k),
                        p["out"], i64(n_out), i32(d), i32(1 if accumulate else 0), None))
    return r["out"]


def sspmm_backward_planned(w4, idx, val, dy, sp_addr, n_cbsr=None, accumulate=False, out_init=None,
                           dxc_align=16, addr_align=16):
    L = load()
    w4 = np.ascontiguousarray(w4, np.int32).reshape(-1)
    sp_addr = np.asarray(sp_addr, np.uint16)
    n_cbsr = sp_addr.shape[0] if n_cbsr is None else n_cbsr
    k, d = sp_addr.shape[1], dy.shape[1]
    ins = {"warp4": Guarded(w4, name="warp4"), "idx": Guarded(np.asarray(idx, np.int32), name="indices", align=4),
           "val": Guarded(np.asarray(val, np.float32), name="values", align=4),
           "dy": Guarded(np.asarray(dy, np.float32), name="dy"),
           "sp_addr": Guarded(sp_addr, name="sp_addr", align=addr_align)}
    out = (Guarded(out_init, name="dxc", align=dxc_align) if out_init is not None
           else Guarded(shape=(n_cbsr, k), dtype=np.float32, name="dxc", align=dxc_align))
    r = _run(L.maxk_sspmm_backward_planned, ins, {"dxc": out},
             lambda p: (p["warp4"], i64(w4.shape[0] // 4), p["idx"], p["val"], p["dy"], p["sp_addr"], i32(k),
                        i32(d), p["dxc"], i64(n_cbsr), i32(1 if accumulate else 0), None))
    return r["dxc"]


def layernorm_forward(x, gamma=None, beta=None, eps=1e-5):
    L = load()
    x = np.ascontiguousarray(x, np.float32)
    n, d = x.shape
    ins = {"x": Guarded(x, name="x")}
    if gamma is not None:
        ins["gamma"] = Guarded(np.asarray(gamma, np.float32), name="gamma")
    if beta is not None:
        ins["beta"] = Guarded(np.asarray(beta, np.float32), name="beta")
    r = _run(L.maxk_layernorm_forward, ins,
             {"y": Guarded(shape=(n, d), dtype=np.float32, name="y"),
              "mean": Guarded(shape=(n,), dtype=np.float32, name="mean", align=4),
              "rstd": Guarded(shape=(n,), dtype=np.float32, name="rstd", align=4)},
             lambda p: (p["x"], i64(n), i32(d), p.get("gamma"), p.get("beta"), ctypes.c_float(eps), p["y"], p["mean"],
                        p["rstd"], None))
    return r["y"], r["mean"], r["rstd"]


def layernorm_backward(g, x, mean, rstd, gamma=None, want_affine=True):
    L = load()
    x = np.ascontiguousarray(x, np.float32)
    n, d = x.shape
    ins = {"g": Guarded(np.asarray(g, np.float32), name="g"), "x": Guarded(x, name="x"),
           "mean": Guarded(np.asarray(mean, np.float32), name="mean", align=4),
           "rstd": Guarded(np.asarray(rstd, np.float32), name="rstd", align=4)}
    if gamma is not None:
        ins["gamma"] = Guarded(np.asarray(gamma, np.float32), name="gamma")
    outs = {"dx": Guarded(shape=(n, d), dtype=np.float32, name="dx")}
    if want_affine:
        outs["dgamma"] = Guarded(shape=(d,), dtype=np.float32, name="dgamma", align=4)
        outs["dbeta"] = Guarded(shape=(d,), dtype=np.float32, name="dbeta", align=4)
    r = _run(L.maxk_layernorm_backward, ins, outs,
             lambda p: (p["g"], p["x"], p["mean"], p["rstd"], p.get("gamma"), i64(n), i32(d), p["dx"], p.get("dgamma"),
                        p.get("dbeta"), None))
    return r["dx"], r.get("dgamma"), r.get("dbeta")


def layernorm_topk(x, gamma, beta, k, eps=1e-5, want_dense=False):
    L = load()
    x = np.ascontiguousarray(x, np.float32)
    n, d = x.shape
    ins = {"x": Guarded(x, name="x")}
    if gamma is not None:
        ins["gamma"] = Guarded(np.asarray(gamma, np.float32), name="gamma")
    if beta is not None:
        ins["beta"] = Guarded(np.asarray(beta, np.float32), name="beta")
    outs = {"sp_data": Guarded(shape=(n, k), dtype=np.float32, name="sp_data", align=4),
            "sp_index": Guarded(shape=(n, k), dtype=np.uint8, name="sp_index", align=1),
            "mean": Guarded(shape=(n,), dtype=np.float32, name="mean", align=4),
            "rstd": Guarded(shape=(n,), dtype=np.float32, name="rstd", align=4)}
    if want_dense:
        outs["y"] = Guarded(shape=(n, d), dtype=np.float32, name="y")
    r = _run(L.maxk_layernorm_topk_cbsr, ins, outs,
             lambda p: (p["x"], i64(n), i32(d), p.get("gamma"), p.get("beta"), ctypes.c_float(eps), i32(k), p["sp_data"],
                        p["sp_index"], p["mean"], p["rstd"], p.get("y"), None))
    return r


def layernorm_backward_cbsr(g_data, sp_index, x, mean, rstd, gamma=None):
    L = load()
    x = np.ascontiguousarray(x, np.float32)
    n, d = x.shape
    k = sp_index.shape[1]
    ins = {"g": Guarded(np.asarray(g_data, np.float32), name="g_data", align=4),
           "si": Guarded(np.asarray(sp_index, np.uint8), name="sp_index", align=1), "x": Guarded(x, name="x"),
           "mean": Guarded(np.asarray(mean, np.float32), name="mean", align=4),
           "rstd": Guarded(np.asarray(rstd, np.float32), name="rstd", align=4)}
    if gamma is not None:
        ins["gamma"] = Guarded(np.asarray(gamma, np.float32), name="gamma")
    outs = {"dx": Guarded(shape=(n, d), dtype=np.float32, name="dx"),
            "dgamma": Guarded(shape=(d,), dtype=np.float32, name="dgamma", align=4),
            "dbeta": Guarded(shape=(d,), dtype=np.float32, name="dbeta", align=4)}
    r = _run(L.maxk_layernorm_backward_cbsr, ins, outs,
             lambda p: (p["g"], p["si"], i32(k), p["x"], p["mean"], p["rstd"], p.get("gamma"), i64(n), i32(d), p["dx"],
                        p["dgamma"], p["dbeta"], None))
    return r["dx"], r["dgamma"], r["dbeta"]
