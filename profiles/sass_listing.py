"""SASS mnemonic counts of the tensor-core conv kernels: `python profiles/sass_listing.py > profiles/r2_sass_conv_tc.txt`
(cuobjdump -sass of the in-tree library; runs without a GPU)."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "from-voxel-to-point_b200", "lib", "libfv2p_b200.so")
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UTMALDG", "UBLKCP", "LDGSTS", "SYNCS", "ELECT", "R2UR", "HMMA",
       "LDS", "STS", "LDG", "STG", "FFMA"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
counts, cur, hmma_total = collections.OrderedDict(), None, 0
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P[0-9T]\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        counts[cur][m.group(1)] += 1
        if m.group(1) == "HMMA":
            hmma_total += 1
print("# SASS mnemonic counts of the tensor-core conv kernels (`cuobjdump -sass libfv2p_b200.so`, sm_100a, round 2)\n")
print("UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM alloc/dealloc, "
      "UBLKCP = cp.async.bulk (weight\nslices, neighbour rows), LDGSTS = cp.async (gathered rows), UTMALDG = "
      "cp.async.bulk.tensor tile::gather4 (optional gather), SYNCS = mbarrier ops.\nTemplate arguments: <fp32 (split-bf16) "
      "path, Cout, packed stages>.  No HMMA (legacy mma.sync) anywhere in the library.\n")
print("| kernel | " + " | ".join(OPS) + " |")
print("|---|" + "---:|" * len(OPS))
for fn, c in counts.items():
    name = demangle(fn)
    if "conv_tc_kernel" not in name:
        continue
    short = re.search(r"conv_tc_kernel<[^>]*>", name)
    label = short.group(0) if short else name
    label = label.replace("(bool)1", "true").replace("(bool)0", "false").replace("(int)", "")
    print("| `%s` | " % label + " | ".join(str(c[o]) for o in OPS) + " |")
print("\nHMMA instructions in the whole library: %d" % hmma_total)
