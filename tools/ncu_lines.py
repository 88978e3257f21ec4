"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line: warp instructions, share,
average active threads, stall samples.  Usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --launch-count 1 | python tools/ncu_lines.py [top]"""
import csv
import sys

top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
cur_file, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit() and r[2] == "-":  # a CUDA line summary row
        d = dict(zip(hdr[4:], r[4:]))
        inst = int(d.get("Instructions Executed", "0") or 0)
        if inst:
            out.append((inst, int(d.get("Thread Instructions Executed", "0") or 0), int(d.get("# Samples", "0") or 0), cur_file, int(r[0]), r[1].strip()[:110]))
tot = sum(o[0] for o in out)
samp = sum(o[2] for o in out)
print("total warp instructions %d, samples %d" % (tot, samp))
for inst, thr, s, f, ln, src in sorted(out, reverse=True)[:top]:
    print("%5.1f%% inst %5.1f%% stall  act %4.1f  %s:%d  %s" % (100.0 * inst / tot, 100.0 * s / max(1, samp), thr / inst, f, ln, src))
