#!/usr/bin/env python
"""Per-kernel SASS evidence of the built library (no GPU needed): counts of the tcgen05 / TMEM / bulk-TMA / mbarrier
mnemonics in every kernel of libendosurf_b200.so (cuobjdump -sass) and the ptxas resource lines of the build logs
(registers, spills, shared memory).  Mnemonics as named in B200_PROFILING.md: UTCHMMA = tcgen05.mma kind::f16,
LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (1-D bulk TMA), UTMALDG = tensor-map TMA, UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops, UTCATOMSWS/UTCALLOC = TMEM allocation.

    python tools/sass_counts.py > profiles/r2_sass_counts.txt
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "endosurf_b200", "libendosurf_b200.so")
MNEMONICS = ["UTCHMMA", "LDTM", "UBLKCP", "UBLKPF", "UTMALDG", "UTCBAR", "SYNCS", "UTCATOMSWS", "HMMA", "FFMA", "MUFU"]


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.strip().splitlines()
        if len(out) == len(names):
            return out
    except Exception:
        pass
    return names


def short(name):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\((?:int|bool)\)", "", name).replace(", ", ",")
    m = re.match(r"([\w:]+)(<[^>]*>)?", name)
    name = m.group(0) if m else name
    if "mlp_chain_kernel<" in name:  # <CHAIN, TANGENT, BWD, STASH, PAIR>
        a = name[name.index("<") + 1:-1].split(",")
        name = (f"es::mlp_chain_kernel<chain={a[0]},{'tangent' if a[1] == '1' else 'plain'},{'bwd' if a[2] == '1' else 'fwd'}"
                f"{',records' if a[3] == '1' else ''}{',pair' if a[4] == '1' else ''}>")
    return name[:78]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = {"name": m.group(1), "n": 0, **{k: 0 for k in MNEMONICS}}
            kernels.append(cur)
            continue
        if cur is None or "/*" not in line:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        cur["n"] += 1
        op = m.group(1).split(".")[0]
        if op in cur:
            cur[op] += 1
    names = demangle([k["name"] for k in kernels])
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels, cuobjdump -sass, built for sm_100a")
    print("# totals: " + ", ".join(f"{m} {sum(k[m] for k in kernels)}" for m in MNEMONICS))
    print(f"{'kernel':80s} {'instr':>7s} " + " ".join(f"{m:>8s}" for m in MNEMONICS))
    for k, nm in sorted(zip(kernels, names), key=lambda t: -t[0]["n"]):
        print(f"{short(nm):80s} {k['n']:7d} " + " ".join(f"{k[m]:8d}" for m in MNEMONICS))
    print("\n# ptxas -v (registers / spill bytes / shared memory) per kernel, from endosurf_b200/csrc/*.o.log")
    csrc = os.path.join(ROOT, "endosurf_b200", "csrc")
    for f in sorted(os.listdir(csrc)):
        if not f.endswith(".o.log"):
            continue
        txt = open(os.path.join(csrc, f)).read()
        ents = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes "
                          r"spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used \d+ barriers)?"
                          r"(?:, (\d+) bytes smem)?", txt)
        if not ents:
            continue
        dn = demangle([e[0] for e in ents])
        print(f"## {f[:-6]}.cu")
        for e, nm in zip(ents, dn):
            print(f"{short(nm):80s} regs {int(e[4]):3d}  stack {int(e[1]):4d} B  spill st/ld {int(e[2]):4d}/{int(e[3]):4d} B"
                  f"  static smem {int(e[5] or 0):6d} B")


if __name__ == "__main__":
    sys.exit(main())
