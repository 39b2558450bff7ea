"""Per-role view of an ncu source page of the warp-specialised eval kernels: splits the SASS at the USETMAXREG markers
(MMA warp | epilogue warps | generator warps) and reports executed instructions, stall samples, and the top instructions.
usage: python scripts/ncu_roles.py file.ncu-rep [top_n]"""
import csv, io, subprocess, sys
f = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 12
out = subprocess.run(['ncu', '-i', f, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
marks = [i for i, r in enumerate(body) if 'USETMAXREG' in r[col['Source']]]
names = ['prologue'] + ['role%d' % k for k in range(len(marks))]
bounds = [0] + marks + [len(body)]
print('markers:', [(i, body[i][col['Source']].strip()) for i in marks])
for k in range(len(bounds) - 1):
    seg = body[bounds[k]:bounds[k + 1]]
    inst = sum(int(r[col['Instructions Executed']] or 0) for r in seg)
    samp = sum(int(r[col['# Samples']] or 0) for r in seg)
    st = {h: sum(int(r[col[h]] or 0) for r in seg) for h in stall_cols}
    tot = sum(st.values()) or 1
    print('== %s: SASS rows %d..%d  warp-inst %.3e  samples %d' % (names[k], bounds[k], bounds[k + 1], inst, samp))
    print('   stalls: ' + ', '.join('%s %.0f%%' % (h[6:], 100 * v / tot) for h, v in sorted(st.items(), key=lambda x: -x[1])[:7]))
    # instruction classes
    cls = {}
    for r in seg:
        op = r[col['Source']].strip().split()
        if not op:
            continue
        o = op[1] if op[0].startswith('@') and len(op) > 1 else op[0]
        o = o.split('.')[0]
        cls[o] = cls.get(o, 0) + int(r[col['Instructions Executed']] or 0)
    print('   ops: ' + ', '.join('%s %.1f%%' % (o, 100 * v / max(inst, 1)) for o, v in sorted(cls.items(), key=lambda x: -x[1])[:14]))
    top = sorted(seg, key=lambda r: -int(r[col['# Samples']] or 0))[:topn]
    for r in top:
        sts = sorted(((h[6:], int(r[col[h]] or 0)) for h in stall_cols), key=lambda x: -x[1])[:2]
        print('   %6s samp %9s exec  %-60s %s' % (r[col['# Samples']], r[col['Instructions Executed']], r[col['Source']].strip()[:60], sts))
