#!/usr/bin/env python3
r"""Compare the SASS of two builds of libmaxk_b200.so kernel by kernel (addresses stripped).

    cuobjdump -sass old.so > old.sass; cuobjdump -sass new.so > new.sass
    python scripts/sass_diff.py old.sass new.sass
    python scripts/sass_diff.py --write-manifest profiles/rNN_sass_manifest.json build.sass   # keep a baseline
    python scripts/sass_diff.py profiles/rNN_sass_manifest.json new.sass

Used to show that a change which only ADDS template variants leaves every kernel that was
validated on the GPU bit-for-bit the same program.  Template parameters appended to FwdCfg with
a default show up as a suffix in the mangled names: pass it with --strip (e.g. ELi0ELi0 for FwdCfg)
or as --map 'REGEX=>REPLACEMENT' pairs (e.g. --map 'sspmm_bwd_kernelI(Lin?\d+)ELb0EEE=>sspmm_bwd_kernelI\1EEE').
"""
import hashlib
import json
import re
import sys


def funcs(path):
    if path.endswith(".json"):  # a manifest written with --write-manifest: kernel -> [n, sha1]
        return {k: tuple(v) for k, v in json.load(open(path)).items()}
    d = _funcs_sass(path)
    return {k: (len(v), hashlib.sha1("\n".join(v).encode()).hexdigest()) for k, v in d.items()}


def _funcs_sass(path):
    d, name, body = {}, None, []
    for ln in open(path):
        m = re.search(r"Function : (\S+)", ln)
        if m:
            if name:
                d[name] = body
            name, body = m.group(1), []
        elif re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
            # addresses stripped, runs of blanks collapsed (cuobjdump pads to the longest line of the FILE)
            body.append(re.sub(r"\s+", " ", re.sub(r"/\*[0-9a-f]+\*/", "", ln)).strip())
    if name:
        d[name] = body
    return d


if __name__ == "__main__":
    strip = ""
    args = sys.argv[1:]
    maps = []  # --map 'REGEX=>REPLACEMENT' (repeatable): rewrite a NEW kernel name into its OLD spelling
    while "--map" in args:
        i = args.index("--map")
        pat, rep = args[i + 1].split("=>", 1)
        maps.append((re.compile(pat), rep))
        del args[i:i + 2]
    if "--strip" in args:
        i = args.index("--strip")
        strip = args[i + 1]
        del args[i:i + 2]
    if "--write-manifest" in args:  # python scripts/sass_diff.py --write-manifest out.json build.sass
        i = args.index("--write-manifest")
        json.dump(funcs(args[i + 2]), open(args[i + 1], "w"), indent=0, sort_keys=True)
        sys.exit(0)
    a, b = funcs(args[0]), funcs(args[1])
    same = diff = 0
    new = []
    for n, body in b.items():
        o = n.replace(strip + "EEEE", "EEEE") if strip and "FwdCfg" in n else n
        for rx, rep in maps:
            o = rx.sub(rep, o)
        if o in a:
            if a[o] == body:
                same += 1
            else:
                diff += 1
                print("DIFF", n, a[o][0], "->", body[0], "instructions")
        else:
            new.append(n)
    def old_name(m):
        o = m.replace(strip + "EEEE", "EEEE") if strip and "FwdCfg" in m else m
        for rx, rep in maps:
            o = rx.sub(rep, o)
        return o
    old_names = {old_name(m) for m in b}
    gone = [n for n in a if n not in b and n not in old_names]
    print(f"identical {same}, changed {diff}, new {len(new)}, removed {len(gone)} (old build: {len(a)} kernels)")
    sys.exit(1 if diff or gone else 0)
