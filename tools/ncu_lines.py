"""Per-source-line stall samples / instruction counts of one kernel from an .ncu-rep (needs -lineinfo + --import-source).
usage: ncu_lines.py report.ncu-rep kernel-regex [top N] [kernel-id filter, e.g. ":::2" = 2nd profiled launch]
The source page comes in several sections per kernel (one per source file + the SASS view); the CUDA-C section with
the most executed instructions is shown, lines sorted by samples."""
import csv, subprocess, sys
from collections import defaultdict
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
sel = ['--kernel-id', sys.argv[4]] if len(sys.argv) > 4 else ['--kernel-name', 'regex:' + rx]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'] + sel,
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, cur, sect = None, None, 0
agg = defaultdict(lambda: [0.0, 0.0, ''])
for r in rows:
    if not r:
        continue
    if r[0] == 'Function Name':
        sect += 1
        continue
    if r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        hdr = {h: i - len(r) for i, h in enumerate(r)}  # from the right: quotes inside source text break the left side
        continue
    if hdr and r[0].isdigit() and len(r) > 10:
        try:
            s, i = float(r[hdr['# Samples']] or 0), float(r[hdr['Instructions Executed']] or 0)
        except (ValueError, KeyError):
            continue
        k = (sect, cur, int(r[0]))
        agg[k][0] += s
        agg[k][1] += i
        agg[k][2] = r[1]
per = defaultdict(float)
for (sc, f, ln), (s, i, src) in agg.items():
    per[sc] += i
best = max(per, key=per.get)
items = [(f, ln, s, i, src) for (sc, f, ln), (s, i, src) in agg.items() if sc == best]
ts, ti = sum(o[2] for o in items) or 1, sum(o[3] for o in items) or 1
print(f'total samples {ts:.0f}  warp instructions {ti:.0f}')
for f, ln, s, i, src in sorted(items, key=lambda o: -o[2])[:top]:
    print(f'{f}:{ln:<4d} samples {s / ts * 100:5.1f}%  inst {i / ti * 100:5.1f}%   {src.strip()[:110]}')
