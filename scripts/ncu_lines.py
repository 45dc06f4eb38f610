"""Aggregate an `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` dump per CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cur = None; hdr = None; out = []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1]; continue
    if len(r) >= 2 and r[0] == 'Function Name': continue
    if len(r) >= 2 and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():      # CUDA line rows carry a line number; SASS rows do not
        ie = int(r[hdr.index('Instructions Executed')] or 0)
        sm = int(r[hdr.index('# Samples')] or 0)
        th = int(r[hdr.index('Thread Instructions Executed')] or 0)
        if ie > 0 or sm > 0: out.append((ie, sm, th, cur.split('/')[-1], r[0], r[1].strip()[:100]))
tot = sum(o[0] for o in out); smp = sum(o[1] for o in out)
print('total inst', tot, 'samples', smp, 'avg active', sum(o[2] for o in out) / max(tot, 1))
for o in sorted(out, reverse=True)[:top]:
    print(f"{o[0]/tot*100:5.2f}% inst {o[1]/max(smp,1)*100:5.2f}% smp act {o[2]/max(o[0],1):4.1f}  {o[3]}:{o[4]}  {o[5]}")
