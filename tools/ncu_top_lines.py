"""Top source lines of an `ncu --page source --csv --print-source sass,cuda` dump (first captured launch)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
secs = [i for i, r in enumerate(rows) if r and r[0] == "File Path"] + [len(rows)]
first_file = rows[secs[0]][1]
agg, samp, src = collections.Counter(), collections.Counter(), {}
for si in range(len(secs) - 1):
    a, b = secs[si], secs[si + 1]
    if si > 0 and rows[a][1] == first_file:
        break  # second captured launch starts here
    f = rows[a][1].split('/')[-1]
    for r in rows[a + 3:b]:
        if len(r) < 9 or not r[0].strip().isdigit() or r[2] != "-":
            continue
        try:
            ie, sm = int(float(r[7] or 0)), int(float(r[6] or 0))
        except ValueError:
            continue
        k = (f, int(r[0]))
        agg[k] += ie; samp[k] += sm; src[k] = r[1].strip()[:120]
tot, ts = sum(agg.values()), sum(samp.values())
print("total warp instructions", tot, "stall samples", ts)
for k, v in agg.most_common(top):
    print("%6.2f%% inst %6.2f%% smp  %s:%d  %s" % (100 * v / tot, 100 * samp[k] / max(ts, 1), k[0], k[1], src[k]))
