"""Per-kernel opcode census of the shipped library (cuobjdump -sass): the instructions that
prove the sm_100a paths are in the binary -- DMMA (FP64 tensor pipe), UTMALDG (tensor-map TMA
loads), UBLKCP (1-D bulk async copies), LDGSTS (cp.async), SYNCS (mbarrier), and that no
library (cuBLAS/cuSOLVER/cuSPARSE) kernel is linked.
usage: python tools/sass_census.py [lib.so] > profiles/sass_census_r02.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "diaglib_b200/libdiaglib_b200.so"
ops = ["DMMA", "DFMA", "UTMALDG", "UBLKCP", "LDGSTS", "SYNCS", "REDUX", "ATOM", "RED", "BAR.SYNC", "SHFL", "MUFU.RSQ", "CCTL"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
counts = collections.OrderedDict()
cur = None
arch = set()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::", "", cur)
        cur = cur.split("(")[0]
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for o in ops:
            if op == o or op.startswith(o + ".") or (o == "BAR.SYNC" and op.startswith("BAR.SYNC")):
                counts[cur][o] += 1
print(f"# SASS opcode census of {lib} (cuobjdump -sass), arch: {', '.join(sorted(arch))}")
deps = subprocess.run(["ldd", lib], capture_output=True, text=True).stdout
libs = [l.split()[0] for l in deps.splitlines() if any(x in l for x in ("cublas", "cusolver", "cusparse", "cudnn", "nccl"))]
print(f"# vendor math / comm libraries linked (ldd): {libs if libs else 'none (NCCL is resolved with dlopen at comm_init)'}")
print(f"# {'kernel':<58}" + "".join(f"{o:>9}" for o in ops) + f"{'instrs':>9}")
tot = collections.Counter()
for k, c in counts.items():
    print(f"  {k[:58]:<58}" + "".join(f"{c[o]:>9}" for o in ops) + f"{c['_total']:>9}")
    tot.update(c)
print(f"  {'TOTAL':<58}" + "".join(f"{tot[o]:>9}" for o in ops) + f"{tot['_total']:>9}")
