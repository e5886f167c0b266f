"""Per-kernel counts of the SASS opcodes that prove (or disprove) a Blackwell-native kernel, from `cuobjdump -sass` of the
built library: UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st: TMEM), UTMALDG (TMA tensor loads), UBLKCP (bulk
copies), HMMA / IMMA (legacy mma.sync pipe), plus the total instruction count.

    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "eddie-wang-hackathon2023_b200", "lib", "libb200_whisper.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "IMMA", "HFMA2", "MUFU"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
    names = iter(demangle)
    cur, counts, total = None, collections.OrderedDict(), collections.Counter()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = next(names)
            cur = re.sub(r"\(.*$", "", cur).replace("void ", "")
            counts.setdefault(cur, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            op = m.group(1)
            total[cur] += 1
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    counts[cur][o] += 1
    print("# cuobjdump -sass eddie-wang-hackathon2023_b200/lib/libb200_whisper.so  (sm_100a) -- opcode counts per kernel")
    print("# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk,")
    print("# HMMA/IMMA = legacy mma.sync pipe.  Kernels without any of these are CUDA-core kernels (glue, filters, packing).")
    hdr = f"{'kernel':<92}" + "".join(f"{o:>8}" for o in OPS) + f"{'total':>8}"
    print(hdr)
    for k, c in counts.items():
        if total[k] == 0:
            continue
        short = k if len(k) <= 90 else k[:87] + "..."
        print(f"{short:<92}" + "".join(f"{c.get(o, 0):>8}" for o in OPS) + f"{total[k]:>8}")
    agg = collections.Counter()
    for c in counts.values():
        agg.update(c)
    print(f"{'ALL KERNELS':<92}" + "".join(f"{agg.get(o, 0):>8}" for o in OPS) + f"{sum(total.values()):>8}")


if __name__ == "__main__":
    sys.exit(main())
