"""cuobjdump -sass opcode histogram per kernel of libflux_b200.so (no GPU needed): the evidence that the hot kernels are
Blackwell-native -- UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA loads, UTCBAR = tcgen05.commit,
no HMMA / HGMMA (legacy tensor paths).  Usage: python profiles/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "flux-generator_b200", "flux", "libflux_b200.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTCOMMA", "UTCMXQMMA", "UTCCP", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "MUFU", "HMMA", "HGMMA",
       "FFMA2", "FADD2", "FMUL2", "F2FP", "LDG", "STG", "LDS", "STS", "SHFL", "BAR", "ELECT")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_x]+)*)", line)
        if m and cur:
            op, mods = m.group(1), m.group(2)
            kernels[cur][op] += 1
            if op.startswith("UTC") and "2CTA" in mods:
                kernels[cur][op + ".2CTA"] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# {os.path.relpath(LIB, ROOT)}: SASS opcode histogram per kernel (cuobjdump -sass, sm_100a)")
    for (name, c), pretty in zip(kernels.items(), demangle):
        total = sum(v for k, v in c.items() if not k.endswith(".2CTA"))
        short = re.sub(r"\(.*", "", pretty)
        keys = [f"{k}={c[k]}" for k in list(KEY) + [k for k in c if k.endswith(".2CTA")] if c.get(k)]
        print(f"{short[:100]:100s} {total:6d} instr | " + " ".join(keys))


if __name__ == "__main__":
    sys.exit(main())
