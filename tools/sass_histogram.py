#!/usr/bin/env python
"""Static issue-slot budget of a kernel: opcode histogram of its SASS, grouped by the pipe that executes it.

    python tools/sass_histogram.py <lib.so|cubin> <substring of the demangled kernel name> [...more substrings]

No GPU needed (cuobjdump reads the embedded sm_100a cubin).  Used to compare epilogue variants before spending GPU time:
the GEMM epilogues at C2 are bound by issue slots / FMA-pipe cycles, so instructions per output element is the figure of
merit (profiles/r01_engineering_log.md, finding 2).
"""
import collections
import re
import subprocess
import sys

PIPES = {
    "fma": ("FFMA", "FMUL", "FADD", "FFMA2", "FMUL2", "FADD2", "IMAD", "HFMA2", "HMUL2", "HADD2"),
    "alu": ("IADD3", "LOP3", "SHF", "PRMT", "FMNMX", "ISETP", "FSETP", "FSET", "SEL", "FSEL", "MOV", "IABS", "LEA", "F2FP", "IADD", "LOP", "SGXT",
            "BMSK", "FLO", "POPC", "VIMNMX", "IMNMX", "FMNMX3"),
    "mufu": ("MUFU",),
    "lsu": ("LDG", "STG", "LDS", "STS", "LDL", "STL", "LD", "ST", "ATOMG", "RED", "ATOMS", "LDC", "LDCU", "LDSM"),
    "tensor/async": ("UTCHMMA", "UTCBAR", "UTMALDG", "UTMAPF", "UTMASTG", "LDTM", "STTM", "SYNCS", "UTCATOMSWS", "UTMACCTL", "ACQBULK", "PREEXIT", "ELECT"),
}


def functions(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, body = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            body[cur].append(m.group(1))
    return body


def main():
    path, pats = sys.argv[1], sys.argv[2:]
    body = functions(path)
    names = list(body)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    for mangled, name in zip(names, dem):
        if pats and not any(p in name for p in pats):
            continue
        h = collections.Counter(body[mangled])
        total = sum(h.values())
        per_pipe = {p: sum(h[o] for o in ops) for p, ops in PIPES.items()}
        other = total - sum(per_pipe.values())
        print(f"{name[:100]}\n  total {total}  " + "  ".join(f"{p} {n}" for p, n in per_pipe.items()) + f"  other {other}")
        print("  " + "  ".join(f"{o} {n}" for o, n in h.most_common(18)))


if __name__ == "__main__":
    main()
