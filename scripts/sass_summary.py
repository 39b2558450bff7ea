"""Count the Blackwell-specific SASS mnemonics per kernel of the built library (evidence for profiles/).
usage: python scripts/sass_summary.py [lib.so] > profiles/r2_sass_counts.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else 'distributedes_b200/libdes_b200.so'
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
pat = ['UTCHMMA.2CTA', 'UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'LDGSTS', 'SYNCS', 'USETMAXREG', 'MUFU', 'FFMA2',
       'HMMA', 'REDG', 'ATOMG', 'MEMBAR', 'NANOSLEEP']
fn, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        continue
    if fn is None:
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m:
        op = m.group(1)
        counts[fn]['_total'] += 1
        for p in pat:
            if op == p or op.startswith(p + '.'):
                counts[fn][p] += 1
                if p == 'UTCHMMA.2CTA':
                    break
demangle = subprocess.run(['c++filt'], input='\n'.join(counts), capture_output=True, text=True).stdout.splitlines()
print('SASS mnemonic counts per kernel of %s (cuobjdump -sass; UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA load)' % lib)
for (fn, c), name in zip(counts.items(), demangle):
    keys = [p for p in pat if c[p]]
    if not any(k in ('UTCHMMA', 'UTCHMMA.2CTA', 'LDTM', 'UTMALDG', 'UBLKCP') for k in keys) and '--all' not in sys.argv:
        continue
    print('%s\n    instructions %d: %s' % (name[:150], c['_total'], ', '.join('%s %d' % (k, c[k]) for k in keys)))
