"""Warp-stall samples per CUDA source line of an ncu --set full --import-source report (needs -lineinfo builds).
usage: ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
lines = []
fname = ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if "# Samples" in r:
        hdr = r
        ix = {h: i for i, h in enumerate(hdr)}
        continue
    if hdr is None or len(r) < len(hdr) or not r[0] or not r[0].isdigit():
        continue
    try:
        n = int(r[ix["# Samples"]])
    except ValueError:
        continue
    st = {h: int(r[ix[h]] or 0) for h in hdr if h.startswith("stall_") and "Not Issued" not in h and r[ix[h]] not in ("", "-")}
    lines.append((n, fname, r[0], r[1].strip(), sorted(st.items(), key=lambda x: -x[1])[:3]))
tot = sum(l[0] for l in lines)
print(f"{path}: {tot} samples")
for n, f, ln, src, st in sorted(lines, key=lambda l: -l[0])[:top]:
    print(f"{n:6d} {100*n/max(tot,1):5.1f}%  {f}:{ln:>4s}  {src[:90]:90s} {st}")
