"""Reduce `ncu -i rep --page source --csv` output to the source lines that collect the most warp-stall samples."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = None
for i, r in enumerate(rows):
    if any("Sampl" in c for c in r) and any(c.strip() in ("Source", "#") or "Source" in c for c in r):
        hdr, start = r, i + 1
        break
if hdr is None:
    print("no header found; first rows:", rows[:3])
    sys.exit(0)
si = next(i for i, c in enumerate(hdr) if "Source" in c)
ci = [i for i, c in enumerate(hdr) if "Sampl" in c and "All" in c] or [i for i, c in enumerate(hdr) if "Sampl" in c]
ci = ci[0]
data = []
for r in rows[start:]:
    if len(r) <= max(si, ci):
        continue
    try:
        n = float(r[ci].replace(",", "") or 0)
    except ValueError:
        continue
    data.append((n, r[0] if si != 0 else "", r[si]))
tot = sum(d[0] for d in data) or 1.0
print(f"# {hdr[ci]} total {tot:.0f}; top lines")
for n, ln, src in sorted(data, key=lambda d: -d[0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 50]:
    print(f"{100 * n / tot:6.2f}%  {n:8.0f}  {ln:>6s}  {src.strip()[:150]}")
