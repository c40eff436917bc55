"""Static counts of the Blackwell tensor-core / TMA / TMEM / mbarrier SASS mnemonics per kernel of the built library:
    python tools/sass_mnemonics.py > profiles/<round>_sass_mnemonics.txt
(cuobjdump -sass polgen-rvc_b200/libpolgen_rvc.so; runs on the GPU-less build host)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "polgen-rvc_b200", "libpolgen_rvc.so")
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UTMALDG", "UBLKCP", "SYNCS", "HMMA", "LDGSTS", "ELECT",
        "ACQBULK", "PREEXIT", "GRIDDEPBAR", "DEPBAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            op = m.group(1)
            for k in KEYS:
                if op.startswith(k):
                    counts[cur][k] += 1
    names = list(counts)
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    for n, d in zip(names, out):
        demangle[n] = re.sub(r"\(.*", "", d.replace("(anonymous namespace)::", ""))
    print("# cuobjdump -sass polgen-rvc_b200/libpolgen_rvc.so : tensor-core / TMA / TMEM mnemonics per kernel (static instruction counts)")
    print("# UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc/dealloc,")
    print("# UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk (1-D bulk copy), SYNCS = mbarrier ops, HMMA = mma.sync,")
    print("# LDGSTS = cp.async, ACQBULK / PREEXIT = griddepcontrol.wait / launch_dependents (programmatic dependent launch)")
    for n in names:
        c = counts[n]
        if not any(c[k] for k in KEYS if k not in ("SYNCS", "DEPBAR")):
            continue
        print(f"{demangle[n][:78]:78s} " + " ".join(f"{k}={c[k]}" for k in sorted(c) if c[k]))


if __name__ == "__main__":
    main()
