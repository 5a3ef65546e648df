#!/usr/bin/env python
"""Write the SASS listings that back the instruction counts quoted in DESIGN.md into profiles/ (no GPU needed):
    python tools/sass_excerpts.py [tag]        -> profiles/<tag>_sass_*.txt
For every kernel of interest: the whole function when it is short, else its hottest loop = the backward branch whose
body holds the most FP64 instructions; plus an opcode histogram of that loop and of the whole function."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "asteroid-spring-sims_b200", "libass_b200.so")
WANT = [
    ("gravity_sym_uniform_r3", r"k_gravity_symILb1ELi3E"),
    ("gravity_sym_uniform_r4", r"k_gravity_symILb1ELi4E"),
    ("gravity_one_sided_ti4", r"k_gravityILi4ELb0E"),
    ("fp64_peak_probe", r"k_dfmaE"),
    ("spring_tiles_kick_uni_pal1", r"k_spring_tilesILb0ELb1ELb1ELb1E"),
    ("ensemble_uni_pal1", r"k_ensembleILb1ELb1E"),
    ("stress", r"8k_stressE"),
]
FP64 = ("DFMA", "DADD", "DMUL", "DSETP", "MUFU.RSQ64H", "MUFU.RCP64H")


def functions(text):
    out, name, buf = {}, None, []
    for ln in text.split("\n"):
        m = re.search(r"Function : (\S+)", ln)
        if m:
            if name:
                out[name] = buf
            name, buf = m.group(1), []
        elif name is not None:
            buf.append(ln)
    if name:
        out[name] = buf
    return out


def instrs(lines):
    res = []
    for ln in lines:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            res.append((int(m.group(1), 16), m.group(2).strip()))
    return res


def opcode(txt):
    t = txt.split()
    if t and t[0].startswith("@"):
        t = t[1:]
    return t[0] if t else "?"


def hottest_loop(ins):
    best = None
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, txt) in enumerate(ins):
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", txt)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr_index:
            j = addr_index[tgt]
            body = ins[j:i + 1]
            n64 = sum(1 for _, t in body if opcode(t).startswith(("DFMA", "DADD", "DMUL")))
            if best is None or n64 > best[0]:
                best = (n64, j, i)
    return best


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    text = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    fns = functions(text)
    outdir = os.path.join(ROOT, "profiles")
    summary = []
    for label, pat in WANT:
        names = [n for n in fns if re.search(pat, n)]
        if not names:
            summary.append("%s: not found (%s)" % (label, pat))
            continue
        name = names[0]
        ins = instrs(fns[name])
        hist_all = collections.Counter(opcode(t) for _, t in ins)
        loop = hottest_loop(ins)
        path = os.path.join(outdir, "%s_sass_%s.txt" % (tag, label))
        with open(path, "w") as f:
            f.write("# %s\n# cuobjdump -sass libass_b200.so (sm_100a), %d instructions in the function\n" % (name, len(ins)))
            f.write("# opcodes (whole function): %s\n" % ", ".join("%s %d" % kv for kv in hist_all.most_common(14)))
            if loop and len(ins) > 400:
                n64, j, i = loop
                body = ins[j:i + 1]
                hist = collections.Counter(opcode(t) for _, t in body)
                f.write("# hottest loop: %d instructions at 0x%x..0x%x; FP64 arithmetic %d (DFMA %d, DADD %d, DMUL %d), MUFU %d, LDS %d, LDG %d, SHFL %d, UBLKCP %d\n" % (
                    len(body), body[0][0], body[-1][0], n64, sum(v for k, v in hist.items() if k.startswith("DFMA")), sum(v for k, v in hist.items() if k.startswith("DADD")),
                    sum(v for k, v in hist.items() if k.startswith("DMUL")), sum(v for k, v in hist.items() if k.startswith("MUFU")),
                    sum(v for k, v in hist.items() if k.startswith("LDS")), sum(v for k, v in hist.items() if k.startswith("LDG")),
                    sum(v for k, v in hist.items() if k.startswith("SHFL")), sum(v for k, v in hist_all.items() if k.startswith("UBLKCP"))))
                for a, t in body:
                    f.write("/*%04x*/ %s\n" % (a, t))
                summary.append("%s: loop of %d instr, %d FP64 arithmetic; function %d instr" % (label, len(body), n64, len(ins)))
            else:
                for a, t in ins:
                    f.write("/*%04x*/ %s\n" % (a, t))
                summary.append("%s: whole function, %d instr" % (label, len(ins)))
    print("\n".join(summary))


if __name__ == "__main__":
    main()
