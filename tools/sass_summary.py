"""Dev tool: instruction mix of every kernel of libslip_lu_b200 from `cuobjdump -sass` (sm_100a),
written as a markdown table (profiles/rNN_sass_summary.md).  Shows which memory / multiply
instructions the hot kernels are made of (LDGSTS = cp.async, IMAD.WIDE, ...) and that none of the
tensor / bulk-copy paths (UTMALDG, UTCMMA/tcgen05, HMMA) is involved.

    python tools/sass_summary.py > profiles/r02_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "slip_lu_b200", "csrc", "_build", "slipcu.o")
KEYS = ["LDGSTS", "LDG", "STG", "LDS", "STS", "IMAD.WIDE", "IMAD.HI", "IMAD", "IADD3", "VIADDMNMX", "VIADD",
        "ISETP", "SHFL", "BAR", "ATOM", "RED", "UTMALDG", "UBLKCP", "UTCMMA", "HMMA", "SYNCS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True, check=True).stdout
    demangle = {}
    names = sorted(set(re.findall(r"Function : (\S+)", sass)))
    if names:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(names, out)) if len(out) == len(names) else {}
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    counts[cur][k] += 1
                    break
    print("# SASS instruction mix per kernel (cuobjdump -sass, sm_100a)\n")
    print("Static instruction counts of the compiled kernels (`python tools/sass_summary.py`). `LDGSTS` is")
    print("`cp.async` (global -> shared without registers); `IMAD.WIDE`/`IMAD`/`IMAD.HI` are the 32-bit")
    print("integer multiplies of the Montgomery arithmetic. No kernel contains a TMA bulk copy (`UTMALDG`/")
    print("`UBLKCP`), an mbarrier wait (`SYNCS`) or a tensor-core instruction (`UTCMMA`, `HMMA`): the path has")
    print("no dense contraction, and the TMA variant of `k_trisolve` was measured slower (see")
    print("`profiles/r01_trisolve_tma_variant_ncu.md`).\n")
    cols = ["total"] + KEYS
    print("| kernel | " + " | ".join(cols) + " |")
    print("|---|" + "---:|" * len(cols))
    for fn, c in counts.items():
        name = demangle.get(fn, fn)
        name = name[:name.rfind("(")] if name.endswith(")") else name
        name = name.replace("(int)", "").replace("(bool)1", "true").replace("(bool)0", "false").replace("void ", "")
        if "k_trisolve" in name and "true, 4>" not in name:
            continue                       # the table keeps the instances with the vector in shared memory and 4 channels per thread (what the bench runs)
        print("| `" + name + "` | " + " | ".join(str(c.get(k, 0)) for k in cols) + " |")


if __name__ == "__main__":
    main()
