"""Per-source-line hot spots of an .ncu-rep captured with --import-source on:
python tools/ncu_lines.py rep [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
stall = sys.argv[3] if len(sys.argv) > 3 else None     # e.g. stall_no_inst: rank lines by it
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None
lines = []
tot_i = tot_s = 0
scol = None
for r in rows:
    if stall and len(r) > 8 and r[0] == 'Line No' and stall in r:
        scol = r.index(stall)
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) < 8 or r[0] in ('Line No', '') or not r[0].isdigit():
        continue
    try:
        samples, inst = int(r[4]), int(r[7])
    except ValueError:
        continue
    if scol is not None:
        try:
            samples = int(r[scol])
        except ValueError:
            samples = 0
    lines.append((inst, samples, cur, int(r[0]), r[1].strip()[:110]))
    tot_i += inst
    tot_s += samples
print('total warp-instructions %d, samples %d' % (tot_i, tot_s))
if stall:
    print('--- by %s samples' % stall)
    for inst, s, f, ln, src in sorted(lines, key=lambda t: -t[1])[:top]:
        print('%5.1f%% inst %5.1f%% %s  %s:%d  %s' % (100. * inst / tot_i, 100. * s / max(1, tot_s), stall, f, ln, src))
print('--- by instructions executed')
for inst, s, f, ln, src in sorted(lines, reverse=True)[:top]:
    print('%5.1f%% inst %5.1f%% stall  %s:%d  %s' % (100. * inst / tot_i, 100. * s / max(1, tot_s), f, ln, src))
byfile = {}
for inst, s, f, ln, src in lines:
    a = byfile.setdefault(f, [0, 0]); a[0] += inst; a[1] += s
print('--- by file')
for f, (i, s) in byfile.items():
    print('%5.1f%% inst %5.1f%% stall %s' % (100. * i / tot_i, 100. * s / max(1, tot_s), f))
