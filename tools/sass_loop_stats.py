#!/usr/bin/env python
"""Static look at the hot loop of a kernel in a cubin / object / .so: instruction mix of the innermost loop that holds most
DFMAs, FP64 instructions per loop trip and register-file reads after reuse-cache hits (a DFMA with three register-file
reads costs 3 issue cycles on B200, with two it costs 2 — tools/micro).  Usage: sass_loop_stats.py <file> <name-regex> [div]"""
import re, subprocess, sys
from collections import Counter


def functions(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        yield f.split("\n")[0], f


def analyse(body_txt, div=1.0):
    ins = []
    for l in body_txt.split("\n"):
        mm = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if mm:
            ins.append((int(mm.group(1), 16), mm.group(2).strip()))
    loops = []
    for addr, t in ins:
        mb = re.search(r"BRA\S*\s+(?:\S+,\s*)*0x([0-9a-f]+)", t)
        if mb and int(mb.group(1), 16) < addr:
            loops.append((int(mb.group(1), 16), addr))
    best = None
    for tgt, addr in loops:
        body = [t for a, t in ins if tgt <= a <= addr]
        nd = sum("DFMA" in t for t in body)
        inner = not any(tgt <= t2 and a2 <= addr and (t2, a2) != (tgt, addr) and
                        sum(1 for a, t in ins if t2 <= a <= a2 and "DFMA" in t) > 8 for t2, a2 in loops)
        if inner and (best is None or nd > best[0]):
            best = (nd, tgt, addr, body)
    if best is None:
        return None
    nd, tgt, addr, body = best
    c = Counter((t.split()[1] if t.startswith("@") else t.split()[0]) for t in body)
    reads = n64 = 0
    prev = {}
    for t in body:
        tt = t.split(None, 1)[1] if t.startswith("@") else t
        op = tt.split()[0]
        if op in ("DFMA", "DADD", "DMUL"):
            n64 += 1
            srcs = [o.strip() for o in tt[len(op):].split(",")[1:]]
            r = 0
            cur = {}
            for slot, o in enumerate(srcs):
                reg = re.match(r"[-|]*(R\d+)", o)
                if not reg:
                    continue
                if prev.get(slot) != reg.group(1):
                    r += 1
                if ".reuse" in o:
                    cur[slot] = reg.group(1)
            reads += r
            prev = cur
        else:
            prev = {}
    other = len(body) - n64
    return dict(loop=(hex(tgt), hex(addr)), n=len(body), mix=dict(c.most_common(14)), fp64=n64, reads=reads, other=other,
                per=dict(fp64=n64 / div, reads=reads / div, other=other / div,
                         cycles_model=(max(2 * n64, reads) + other) / div))


if __name__ == "__main__":
    path, pat = sys.argv[1], sys.argv[2]
    div = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    for name, f in functions(path):
        if re.search(pat, name):
            r = analyse(f, div)
            print(name[:150])
            print("  ", r)
