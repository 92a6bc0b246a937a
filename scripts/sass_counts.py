"""Evidence for the tcgen05 / TMEM / bulk-copy claims: per-kernel counts of the Blackwell SASS mnemonics in the built
library (cuobjdump -sass), written to profiles/<tag>_sass_counts.txt.  Runs on the CPU box."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "paif_b200", "libpaif_b200.so")
WANT = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTCBAR", "SYNCS", "HMMA", "FFMA", "SHFL", "LDGSTS"]


def main(tag):
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for w in WANT:
                if op.startswith(w):
                    kernels[cur][w] += 1
    out = ["SASS mnemonic counts per kernel of paif_b200/libpaif_b200.so (cuobjdump -sass; sm_100a).",
           "UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops.",
           "", "%-110s %7s " % ("kernel", "instrs") + " ".join("%7s" % w for w in WANT)]
    tot = collections.Counter()
    for k, c in kernels.items():
        if not any(c[w] for w in ("UTCHMMA", "LDTM", "STTM", "UBLKCP")):
            continue
        name = re.sub(r"\(.*", "", k)[:108]
        out.append("%-110s %7d " % (name, c["_total"]) + " ".join("%7d" % c[w] for w in WANT))
        tot.update(c)
    out.append("%-110s %7d " % ("TOTAL (kernels using tcgen05 / TMEM / bulk copies)", tot["_total"]) + " ".join("%7d" % tot[w] for w in WANT))
    out.append("")
    out.append("kernels without tensor-core / bulk-copy instructions (direct FFMA / pointwise / marching kernels): %d"
               % sum(1 for c in kernels.values() if not any(c[w] for w in ("UTCHMMA", "LDTM", "STTM", "UBLKCP"))))
    path = os.path.join(ROOT, "profiles", "%s_sass_counts.txt" % tag)
    open(path, "w").write("\n".join(out) + "\n")
    print(path)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r2")
