#!/usr/bin/env python
"""Per-kernel SASS instruction-count summary of libpsg_b200.so (evidence that the contraction kernels are tcgen05 / TMA
native and that FPS uses packed fp32x2 + REDUX).  Runs without a GPU:

    python tools/sass_summary.py > profiles/r2_sass_summary.txt

Mnemonics (B200_PROFILING.md): UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tile load,
UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier, REDUX = warp reduction, FADD2 / FMUL2 / FFMA2 = fp32x2,
HMMA / IMMA = legacy mma.sync (none expected)."""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(REPO, "pointsecguard_b200", "libpsg_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "REDUX", "FADD2", "FMUL2", "FFMA2",
        "FFMA", "LDG", "STG", "LDS", "STS", "ATOM", "RED", "HMMA", "IMMA", "BAR"]


def strip_params(name):
    """drop the function parameter list (the first '(' outside template brackets), keep the template arguments"""
    depth = 0
    for i, ch in enumerate(name):
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            return name[:i]
    return name


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    demangle = {}
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            per[cur]["_total"] += 1
            for k in KEYS:
                if op == k or (k in ("LDG", "STG", "LDS", "STS", "ATOM", "RED", "BAR", "SYNCS", "REDUX") and op.startswith(k)):
                    per[cur][k] += 1
    names = list(per)
    try:
        dm = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        demangle = dict(zip(names, dm))
    except Exception:
        pass
    print(f"# {os.path.relpath(SO, REPO)}: {len(per)} kernels; columns = instruction counts in the sm_100a SASS")
    print("kernel\ttotal\t" + "\t".join(KEYS))
    tot = collections.Counter()
    for n, c in per.items():
        short = strip_params(re.sub(r"^void ", "", demangle.get(n, n)))
        print(short + "\t" + str(c["_total"]) + "\t" + "\t".join(str(c[k]) for k in KEYS))
        tot.update(c)
    print("ALL\t" + str(tot["_total"]) + "\t" + "\t".join(str(tot[k]) for k in KEYS))


if __name__ == "__main__":
    main()
