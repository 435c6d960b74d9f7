"""Worst figures of a scripts/debug_parity.py log: ours vs fp32 oracle, ours vs fp64 twin, oracle vs fp64 twin."""
import re, sys
rows = []
for l in open(sys.argv[1]):
    m = re.match(r'(\S+)\s+vs fp32 oracle (\S+) (\S+) \| vs fp64 (\S+) (\S+) \| oracle32 vs fp64 (\S+) (\S+)', l)
    if m:
        rows.append((m.group(1),) + tuple(float(x) for x in m.groups()[1:]))
if not rows:
    print("no rows in", sys.argv[1]); sys.exit(0)
w32 = max(rows, key=lambda r: max(r[1], r[2])); w64 = max(rows, key=lambda r: max(r[3], r[4])); wr = max(rows, key=lambda r: max(r[5], r[6]))
ratio = max(rows, key=lambda r: max(r[3], r[4]) / max(r[5], r[6]))
print(f"{sys.argv[1].split('/')[-1]}: worst ours-vs-fp32 {max(w32[1], w32[2]):.2e} ({w32[0]}) | ours-vs-fp64 {max(w64[3], w64[4]):.2e} ({w64[0]}) | "
      f"oracle32-vs-fp64 {max(wr[5], wr[6]):.2e} ({wr[0]}) | worst ratio {max(ratio[3], ratio[4]) / max(ratio[5], ratio[6]):.1f} ({ratio[0]}); "
      f"median ours-vs-fp64 fro {sorted(r[4] for r in rows)[len(rows) // 2]:.2e}")
