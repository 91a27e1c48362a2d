#!/usr/bin/env python3
"""Top source lines of an `ncu --page source --csv --print-source cuda,sass` dump by stall samples / instructions.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv; python tools/ncu_lines.py x.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], newline='')))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out, fname, hdr = [], '', None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr and r[0].isdigit():
        si, ii = hdr.index('# Samples'), hdr.index('Instructions Executed')
        try: out.append((int(r[si] or 0), int(r[ii] or 0), fname, int(r[0]), r[1].strip(), r, hdr))
        except ValueError: pass
ts, ti = sum(o[0] for o in out) or 1, sum(o[1] for o in out) or 1
print(f"total samples {ts}  warp instructions {ti}")
stall_cols = [c for c in out[0][6] if c.startswith('stall_') and 'Not Issued' not in c]
for s, i, f, ln, src, r, h in sorted(out, key=lambda o: -o[0])[:N]:
    st = sorted(((int(r[h.index(c)] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{s*100/ts:5.1f}% smp {i*100/ti:5.1f}% ins {f}:{ln:<4} {' '.join(f'{n}:{v}' for v, n in st):28} {src[:110]}")
