"""Per-source-line stall samples / instruction counts of one kernel from an .ncu-rep (needs -lineinfo + --import-source).
usage: ncu_lines.py report.ncu-rep kernel-regex [top N] [launch index among the matches]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv', '--kernel-name',
                      'regex:' + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur_file, hdr, out, seen_fn = None, None, [], 0
for r in rows:
    if not r:
        continue
    if r[0] == 'Function Name':
        seen_fn += 1
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        hdr = {h: i - len(r) for i, h in enumerate(r)}  # from the right: quotes inside source text break the left side
        continue
    if hdr and r[0].isdigit() and len(r) >= -hdr['# Samples']:
        try:
            out.append((cur_file, int(r[0]), r[1], float(r[hdr['# Samples']] or 0),
                        float(r[hdr['Instructions Executed']] or 0), seen_fn))
        except ValueError:
            pass
# several launches of the same kernel: metrics repeat per function block; keep the first block only
blocks = sorted({o[5] for o in out})
pick = blocks[min(which, len(blocks) - 1)] if blocks else 0
first = {}
for f, ln, src, s, i, b in out:
    if b == pick:
        first.setdefault((f, ln), (f, ln, src, s, i))
out = list(first.values())
ts, ti = sum(o[3] for o in out) or 1, sum(o[4] for o in out) or 1
print(f'total samples {ts:.0f}  warp instructions {ti:.0f}')
for f, ln, src, s, i in sorted(out, key=lambda o: -o[3])[:top]:
    print(f'{f}:{ln:<4d} samples {s / ts * 100:5.1f}%  inst {i / ti * 100:5.1f}%   {src.strip()[:110]}')
