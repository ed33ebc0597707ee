"""Per-source-line totals (instructions executed, samples) from an .ncu-rep captured with --import-source on."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
hdr = None
lines = []
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
    elif r[0] == 'Line No':
        hdr = r
        ix = {h: i for i, h in enumerate(hdr)}
    elif hdr and r[0] not in ('', 'Function Name') and len(r) > 8:
        try:
            lines.append((cur_file, int(r[0]), r[1], int(r[ix['Instructions Executed']] or 0), int(r[ix['# Samples']] or 0)))
        except ValueError:
            pass
tot = sum(l[3] for l in lines)
tots = sum(l[4] for l in lines)
print('total instructions', tot, 'samples', tots)
for f, n, src, ex, smp in sorted(lines, key=lambda l: -l[3])[:top]:
    print('%5.1f%% inst %5.1f%% smp  %s:%d  %s' % (100.0 * ex / tot, 100.0 * smp / max(1, tots), f, n, src.strip()[:110]))
