"""SASS evidence per object file of libpfn_b200.so: counts of the mnemonics that prove the Blackwell-native paths
(B200_PROFILING.md): UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG (TMA tensor copies), UBLKCP
(cp.async.bulk), UBLKPF (bulk L2 prefetch), LDGSTS (cp.async), SYNCS (mbarrier), UTCBAR (tcgen05.commit), HMMA (legacy
mma.sync: must be absent).  Writes profiles/r2_sass_<object>.txt (per-kernel table) and prints a summary.

    python scripts/sass_table.py            # after `python -m poweflownet_b200.build`
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "poweflownet_b200", "lib", "obj")
OUT = os.path.join(ROOT, "profiles")
PATTERNS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF", "LDGSTS", "SYNCS", "UTCBAR", "HMMA", "ACQBULK",
            "LDG", "STG", "LDS", "STS", "FFMA", "BAR"]
summary = []
for obj in sorted(os.listdir(OBJ)):
    if not obj.endswith(".o"):
        continue
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
    kernels, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"pfn::\(anonymous namespace\)::", "", cur)
            cur = re.sub(r"\(.*", "", cur)
            kernels[cur] = {p: 0 for p in PATTERNS}
            kernels[cur]["instructions"] = 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        kernels[cur]["instructions"] += 1
        for p in PATTERNS:
            if op.startswith(p) and not (p == "BAR" and not op.startswith("BAR.")) and not (p == "LDG" and op.startswith("LDGSTS")):
                kernels[cur][p] += 1
    cols = ["instructions"] + PATTERNS
    with open(os.path.join(OUT, f"r2_sass_{obj[:-2]}.txt"), "w") as fh:
        fh.write(f"# cuobjdump -sass poweflownet_b200/lib/obj/{obj}: instruction counts per kernel (sm_100a)\n")
        fh.write("kernel".ljust(64) + "".join(c.rjust(13) for c in cols) + "\n")
        for k, v in kernels.items():
            fh.write(k[:63].ljust(64) + "".join(str(v[c]).rjust(13) for c in cols) + "\n")
    tot = {c: sum(v[c] for v in kernels.values()) for c in cols}
    summary.append((obj, len(kernels), tot))
print("object".ljust(22) + "kernels".rjust(8) + "".join(c.rjust(9) for c in ["UTC*MMA", "LDTM", "UTMALDG", "UBLKCP", "UBLKPF", "LDGSTS", "SYNCS", "UTCBAR", "HMMA"]))
for obj, nk, t in summary:
    print(obj.ljust(22) + str(nk).rjust(8) + "".join(str(x).rjust(9) for x in [t["UTCHMMA"] + t["UTCQMMA"] + t["UTCIMMA"], t["LDTM"], t["UTMALDG"], t["UBLKCP"], t["UBLKPF"],
                                                                                 t["LDGSTS"], t["SYNCS"], t["UTCBAR"], t["HMMA"]]))
