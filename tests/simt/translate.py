#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: rewrite the CUDA sources of spgemm-new_b200/csrc into C++ that g++ can
compile against tests/simt/include/cuda_runtime.h (the SIMT test interpreter).

Only what a C++ compiler cannot parse is touched; the kernel bodies stay as they are:

  * `kernel<<<grid, block, smem, stream>>>(args);`
        -> `::simt::launch((const void *)(&kernel), ::simt::LaunchCfg(grid, block, smem, stream),
                           [&]() { kernel(args); });`
  * `extern __shared__ <attrs> T name[];`
        -> `T *name = reinterpret_cast<T *>(::simt::dynamic_smem());`
  * `__noinline__` -> SIMT_NOINLINE (the bare word collides with GCC's attribute spelling)
  * `asm [volatile](...)` statements: the handful of PTX instructions the single-GPU kernels
    execute get a C++ stand-in (table PTX below); every other one (TMA, mbarrier, fences of the
    NVLink path, which the interpreter cannot run) becomes a call that fails the launch if it is
    ever reached.
"""
import os
import re
import sys


def _balanced(text, start, open_ch="(", close_ch=")"):
    """index just past the bracket that closes text[start] (which must be open_ch); skips
    string and character literals."""
    assert text[start] == open_ch, (text[start - 20:start + 20], open_ch)
    depth, i, n = 0, start, len(text)
    while i < n:
        c = text[i]
        if c == '"':
            if text.startswith('R"(', i - 1):
                i = text.index(')"', i) + 2
                continue
            i += 1
            while text[i] != '"':
                i += 2 if text[i] == "\\" else 1
        elif c == "'":
            i += 1
            while text[i] != "'":
                i += 2 if text[i] == "\\" else 1
        elif c == open_ch:
            depth += 1
        elif c == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise ValueError("unbalanced brackets")


def _split_top(s, sep):
    """split on `sep` at bracket depth 0, outside string literals"""
    out, depth, cur, i = [], 0, [], 0
    while i < len(s):
        c = s[i]
        if c == '"':
            j = i + 1
            while s[j] != '"':
                j += 2 if s[j] == "\\" else 1
            cur.append(s[i:j + 1])
            i = j + 1
            continue
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == sep and depth == 0:
            out.append("".join(cur))
            cur = []
        else:
            cur.append(c)
        i += 1
    out.append("".join(cur))
    return out


# ---- inline PTX ---------------------------------------------------------------------------------
# pattern on the PTX text -> C++ with {0}, {1}, ... = the operand expressions in PTX order
PTX = [
    (r"^red\.relaxed\.gpu\.global\.add\.v4\.f32 ",
     "{{ float *p_ = (float *)({0}); p_[0] += ({1}); p_[1] += ({2}); p_[2] += ({3}); p_[3] += ({4}); }}"),
    (r"^red\.relaxed\.gpu\.global\.add\.f32 ", "{{ *(float *)({0}) += ({1}); }}"),
    (r"^rcp\.approx\.ftz\.f32 ", "{{ ({0}) = 1.0f / ({1}); }}"),
    (r"^redux\.sync\.max\.f32 %0, %1, 0xffffffff;", "{{ ({0}) = ::simt_redux_f32(({1}), true); }}"),
    (r"^redux\.sync\.min\.f32 %0, %1, 0xffffffff;", "{{ ({0}) = ::simt_redux_f32(({1}), false); }}"),
    (r"shfl\.sync\.up\.b32 r\|p, %0, %1, 0, 0xffffffff; @p add\.u32 %0, %0, r;",
     "{{ unsigned n_ = __shfl_up_sync(0xffffffffu, ({0}), (unsigned)({1})); "
     "if ((int)(threadIdx.x & 31u) >= (int)({1})) ({0}) += n_; }}"),
    (r"setp\.geu\.f32 p, %1, %2; @p add\.s32 %0, %0, 1;", "{{ if (!(({1}) < ({2}))) ({0}) += 1; }}"),
    (r"setp\.geu\.f32 p, %1, %2; @p st\.shared\.f32 \[%0\], %1; @p st\.shared\.u32 \[%0\+%4\], %3;",
     "{{ if (!(({1}) < ({2}))) {{ *(float *)::simt::smem_resolve({0}) = ({1}); "
     "*(unsigned *)::simt::smem_resolve(({0}) + ({4})) = ({3}); ({0}) += 4; }} }}"),
    (r"^ld\.acquire\.(sys|gpu)\.global\.u32 ", "{{ ({0}) = *(const volatile unsigned *)({1}); }}"),
    (r"^st\.release\.(sys|gpu)\.global\.u32 ", "{{ *(volatile unsigned *)({0}) = ({1}); }}"),
    (r"^ld\.global\.L1::no_allocate\.v4\.u32 ",
     "{{ const unsigned *p_ = (const unsigned *)({4}); ({0}) = p_[0]; ({1}) = p_[1]; ({2}) = p_[2]; ({3}) = p_[3]; }}"),
    (r"^cp\.async\.cg\.shared\.global ", "::simt::cp_async(::simt::smem_resolve({0}), (const void *)({1}), 16)"),
    # TMA unit: mbarrier + bulk copies (NVLink push of the peer path)
    (r"^mbarrier\.init\.shared::cta\.b64 \[%0\], 1;", "::simt::mbar_init({0}, 1)"),
    (r"^fence\.(mbarrier_init|proxy\.async)", "(void)0"),
    (r"^mbarrier\.arrive\.expect_tx\.shared::cta\.b64 _, ", "::simt::mbar_expect_tx({0}, {1})"),
    (r"^cp\.async\.bulk\.shared::cluster\.global\.mbarrier::complete_tx::bytes ",
     "::simt::bulk_load({0}, (const void *)({1}), {2}, {3})"),
    (r"mbarrier\.try_wait\.parity\.shared::cta\.b64 p, \[%0\], %1;", "::simt::mbar_wait({0}, {1})"),
    (r"^cp\.async\.bulk\.global\.shared::cta\.bulk_group ", "::simt::bulk_store((void *)({0}), {1}, {2})"),
    (r"^cp\.async\.bulk\.commit_group", "(void)0"),
    (r"^cp\.async\.bulk\.wait_group", "::simt::bulk_wait()"),
    (r"^cp\.async\.commit_group", "::simt::cp_async_commit()"),
    (r"^cp\.async\.wait_group ", "::simt::cp_async_wait({0})"),
    (r"^atom\.acq_rel\.gpu\.global\.add\.u32 ",
     "{{ unsigned *p_ = (unsigned *)({1}); ({0}) = *p_; *p_ += 1; }}"),
]


def _asm_statement(stmt):
    """stmt: the text between the parentheses of asm(...)"""
    parts = _split_top(stmt, ":")
    ptx = "".join(re.findall(r'"((?:[^"\\]|\\.)*)"', parts[0]))
    ptx_flat = re.sub(r"\s+", " ", ptx.replace("\\n", " ").replace("\\t", " ")).strip()
    operands = []
    for sec in parts[1:3]:  # outputs, inputs
        for op in _split_top(sec, ","):
            op = op.strip()
            if not op:
                continue
            m = re.match(r'"[^"]*"\s*\(', op)
            assert m, ("operand", op)
            operands.append(op[m.end():op.rindex(")")].strip())
    for pat, repl in PTX:
        if re.search(pat, ptx_flat):
            return repl.format(*operands)
    esc = ptx_flat.replace("\\", "\\\\").replace('"', '\\"')
    return '::simt::unsupported_asm("%s")' % esc


def _rewrite_asm(text):
    out, i = [], 0
    rx = re.compile(r"\basm\s*(volatile\s*)?\(")
    while True:
        m = rx.search(text, i)
        if not m:
            out.append(text[i:])
            break
        out.append(text[i:m.start()])
        end = _balanced(text, m.end() - 1)
        out.append(_asm_statement(text[m.end():end - 1]))
        i = end  # the statement's own ';' stays
    return "".join(out)


# ---- launches -----------------------------------------------------------------------------------
def _rewrite_launches(text):
    out, i = [], 0
    while True:
        j = text.find("<<<", i)
        if j < 0:
            out.append(text[i:])
            break
        m = re.search(r"([A-Za-z_][\w:]*(?:<[^<>;(){}]*>)?)\s*$", text[:j])
        assert m, "kernel name before <<< not found: " + text[max(0, j - 80):j]
        kernel = m.group(1)
        k = text.index(">>>", j)
        cfg = text[j + 3:k]
        p = text.index("(", k)
        assert text[k + 3:p].strip() == "", text[k:p + 1]
        e = _balanced(text, p)
        args = text[p + 1:e - 1]
        out.append(text[i:m.start(1)])
        out.append("::simt::launch((const void *)(&%s), ::simt::LaunchCfg(%s), [&]() { %s(%s); })"
                   % (kernel, cfg, kernel, args))
        i = e
    return "".join(out)


_SHARED = re.compile(r"extern\s+__shared__\s+((?:__align__\(\d+\)\s+)?)([\w\s]+?)\s*\b(\w+)\[\];")


def translate(text):
    text = _rewrite_asm(text)
    text = _rewrite_launches(text)
    text = _SHARED.sub(lambda m: "%s *%s = reinterpret_cast<%s *>(::simt::dynamic_smem());"
                       % (m.group(2), m.group(3), m.group(2)), text)
    text = re.sub(r"\b__noinline__\b", "SIMT_NOINLINE", text)
    assert "<<<" not in text and "extern __shared__" not in text
    return text


# tcgen05 / TMA tensor maps / TMEM have no stand-in here: these files run on a B200 only (their C entry points are
# stubbed by tensor_core_stubs.cpp so that the library still exports the whole header)
TENSOR_CORE_ONLY = {"linear_topk.cu", "linear_wgrad.cu", "tc_common.cuh"}


def translate_tree(src_dir, dst_dir):
    os.makedirs(dst_dir, exist_ok=True)
    done = []
    for name in sorted(os.listdir(src_dir)):
        if not name.endswith((".cu", ".cuh")) or name in TENSOR_CORE_ONLY:
            continue
        with open(os.path.join(src_dir, name)) as f:
            text = translate(f.read())
        # .cuh keep their names (they are #included by name); .cu become .cpp
        out = name if name.endswith(".cuh") else name[:-3] + ".cpp"
        path = os.path.join(dst_dir, out)
        head = "// GENERATED by tests/simt/translate.py from spgemm-new_b200/csrc/%s -- do not edit\n" % name
        old = open(path).read() if os.path.exists(path) else None
        if old != head + text:
            with open(path, "w") as f:
                f.write(head + text)
        done.append(path)
    return done


if __name__ == "__main__":
    for p in translate_tree(sys.argv[1], sys.argv[2]):
        print(p)
