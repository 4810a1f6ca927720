"""Per-kernel SASS histogram of libnlps_b200.so: counts of the Blackwell-native mnemonics (tcgen05 MMA / TMEM load-store /
TMA) and of the legacy tensor-core instruction, per kernel function.  Proof that the hot kernels are sm_100a-native.

    python tools/sass_histogram.py > profiles/r02_sass_histogram.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "nlpscratch_b200", "libnlps_b200.so")
WANT = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "SYNCS"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if not m:
        continue
    op = m.group(1)
    base = op.split(".")[0]
    counts[cur]["_total"] += 1
    if base == "UTCHMMA":
        counts[cur]["UTCHMMA"] += 1
        if ".2CTA" in op:
            counts[cur]["UTCHMMA.2CTA"] += 1
    elif base in ("LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS"):
        counts[cur][base] += 1
    elif base == "HMMA":
        counts[cur]["HMMA"] += 1


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    out = [n.replace("nlps::(anonymous namespace)::", "").replace("(anonymous namespace)::", "").replace("nlps::", "") for n in out]
    return [re.sub(r"^void ", "", re.sub(r"\(.*", "", n)) for n in out]


names = list(counts)
short = demangle(names)
print("# SASS histogram of `nlpscratch_b200/libnlps_b200.so` (%s only; `cuobjdump -sass`)\n" % ", ".join(arch))
print("UTCHMMA = tcgen05.mma (`.2CTA` = cta_group::2), LDTM / STTM = tcgen05.ld / tcgen05.st (tensor memory), UTMALDG / UTMASTG = TMA "
      "load / store (cp.async.bulk.tensor), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = legacy mma.sync.\n")
print("| kernel | instructions | " + " | ".join(WANT) + " |\n|---|---:|" + "---:|" * len(WANT))
tot = collections.Counter()
for n, s in zip(names, short):
    c = counts[n]
    if not any(c[w] for w in WANT if w != "SYNCS"):
        continue
    print("| `%s` | %d | " % (s, c["_total"]) + " | ".join(str(c[w]) if c[w] else "" for w in WANT) + " |")
    tot.update(c)
print("| **all kernels with tensor-core / TMA instructions** | %d | " % tot["_total"] + " | ".join(str(tot[w]) for w in WANT) + " |")
plain = [s for n, s in zip(names, short) if not any(counts[n][w] for w in WANT if w != "SYNCS")]
print("\n%d further kernels (bandwidth-bound element-wise / reduction / indexing kernels, the CUDA-core FP32 GEMM and attention of "
      "the `fp32` mode) use none of these instructions." % len(plain))
