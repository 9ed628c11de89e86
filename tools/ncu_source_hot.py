"""Hot spots of an `ncu --page source --csv --print-source sass` export: per kernel launch, the SASS instructions with the
most stall samples, the executed-instruction total and the stall-reason totals.
    ncu -i x.ncu-rep --page source --csv --print-source sass > x_src.csv;  python tools/ncu_source_hot.py x_src.csv [topN]"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path, newline="")))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
for b in blocks:
    h = {n: i for i, n in enumerate(b["hdr"])}
    stall_cols = [n for n in b["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
    tot = {n: 0 for n in stall_cols}
    samples = inst = 0
    recs = []
    for idx, r in enumerate(b["rows"]):
        s = int(r[h["# Samples"]] or 0)
        e = int(r[h["Instructions Executed"]] or 0)
        samples += s
        inst += e
        for n in stall_cols:
            tot[n] += int(r[h[n]] or 0)
        top_stall = max(stall_cols, key=lambda n: int(r[h[n]] or 0))
        recs.append((s, idx, r[h["Source"]].strip(), e, top_stall, int(r[h[top_stall]] or 0),
                     r[h["L1 Wavefronts Shared"]], r[h["L1 Wavefronts Shared Ideal"]]))
    print("==", b["name"][:90])
    print("   samples %d   warp instructions executed %d   SASS lines %d" % (samples, inst, len(recs)))
    print("   stall totals:", ", ".join("%s %.1f%%" % (n[6:], 100.0 * v / max(samples, 1))
                                        for n, v in sorted(tot.items(), key=lambda kv: -kv[1])[:9]))
    for s, idx, src, e, ts, tv, wf, wfi in sorted(recs, reverse=True)[:top]:
        print("   %5.2f%%  line %4d  exec %9d  %-18s %5d  smem wf %s/%s  %s" % (100.0 * s / max(samples, 1), idx, e, ts[6:], tv, wf, wfi, src[:70]))
