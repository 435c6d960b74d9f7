"""Per-source-line totals from an ncu report compiled with -lineinfo: executed warp instructions and stall samples,
hottest lines first.  python scripts/ncu_lines.py report.ncu-rep [top_n]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lines, cur_file, hdr = {}, None, None
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ia, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= ia or r[0] == "":
        continue  # SASS rows repeat what the CUDA-line row already totals
    try:
        n, s = int(r[ia]), int(r[isamp])
    except ValueError:
        continue
    key = (cur_file, r[0])
    a = lines.setdefault(key, [0, 0, r[1].strip()[:120]])
    a[0] += n
    a[1] += s
tot_i, tot_s = sum(v[0] for v in lines.values()), sum(v[1] for v in lines.values())
print(f"total warp instructions {tot_i}, stall samples {tot_s}")
for (f, ln), (n, s, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{ln:>4s} inst {n:9d} ({100.0 * n / max(tot_i, 1):5.1f}%) samples {s:6d} ({100.0 * s / max(tot_s, 1):5.1f}%)  {src}")
