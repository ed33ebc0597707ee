"""Per-SASS-instruction view of an .ncu-rep source page: executed counts, samples, shared-memory wavefronts."""
import csv, subprocess, sys
rep = sys.argv[1]
minexec = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
tot_inst = tot_samp = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    ex = int(r[ix['Instructions Executed']] or 0)
    tot_inst += ex
    tot_samp += int(r[ix['# Samples']] or 0)
print('total warp instructions', tot_inst, 'samples', tot_samp)
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    ex = int(r[ix['Instructions Executed']] or 0)
    if ex < minexec:
        continue
    stalls = {k[6:]: int(r[ix[k]] or 0) for k in hdr if k.startswith('stall_') and 'Not Issued' not in k}
    top = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
    print('%6s %8d smp %5s wf %7s/%-7s %-70s %s' % (r[ix['Address']][-5:], ex, r[ix['# Samples']], r[ix['L1 Wavefronts Shared']],
          r[ix['L1 Wavefronts Shared Ideal']], r[ix['Source']][:70], ' '.join('%s:%d' % t for t in top if t[1])))
