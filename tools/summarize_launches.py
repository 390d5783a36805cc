"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares."""
import collections, csv, re, sys
path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0      # launches to skip (warm-up)
lines = [l for l in open(path) if not l.startswith("==")]
rows = []
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    rows.append((re.sub(r"\(.*", "", row["Kernel Name"]).replace("<unnamed>::", "").replace("ut2::", ""), v))
rows = rows[skip:]
tot, cnt = collections.defaultdict(float), collections.Counter()
for n, v in rows:
    tot[n] += v; cnt[n] += 1
T = sum(tot.values())
print(f"# {path}: {len(rows)} launches after skipping {skip}; sum of durations {T/1e3:.2f} ms (cold-cache, serialised: compare shares)")
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:40]:
    print(f"{100*v/T:6.2f}%  {v/1e3:9.3f} ms  n={cnt[k]:5d}  avg={v/cnt[k]:8.1f} us  {k[:100]}")
