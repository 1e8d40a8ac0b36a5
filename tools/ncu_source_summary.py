"""Summarise an `ncu --page source --csv` dump: stall totals and the hottest SASS lines.
usage: ncu -i rep.ncu-rep --page source --csv > src.csv; python tools/ncu_source_summary.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
# (a dump may hold several views -- SASS, then source -- each with its own header: take the first)
end = next((i for i in range(hdr_i + 1, len(rows)) if rows[i] and rows[i][0] in ("Address", "Kernel Name", "#")), len(rows))
body = [r for r in rows[hdr_i + 1:end] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: 0 for s in stalls}
samples = 0
for r in body:
    samples += int(r[col["# Samples"]] or 0)
    for s in stalls:
        tot[s] += int(r[col[s]] or 0)
print("total samples", samples)
for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print("  %-24s %8d  %5.1f%%" % (s, v, 100.0 * v / max(samples, 1)))
print("hottest instructions:")
for r in sorted(body, key=lambda r: -int(r[col["# Samples"]] or 0))[:top]:
    st = sorted(((int(r[col[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print("  %6s %5.1f%%  %-70s %s" % (r[col["# Samples"]], 100.0 * int(r[col["# Samples"]]) / samples,
                                       r[col["Source"]][:70], " ".join("%s=%d" % (s[6:], v) for v, s in st)))
