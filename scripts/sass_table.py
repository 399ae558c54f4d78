"""profiles/sass_opcodes.md: per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use
(cuobjdump -sass of the shipped library).   python scripts/sass_table.py > profiles/sass_opcodes.md"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, 'dposer_b200', 'lib', 'libdposer_b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
ops = ['UTCHMMA', 'UTCHMMA.2CTA', 'UTCBAR', 'LDTM', 'UTMALDG', 'UTMASTG', 'SYNCS', 'UCGABAR', 'FFMA', 'STG', 'ATOM', 'RED']
rows, cur, counts = [], None, None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        if cur:
            rows.append((cur, counts))
        cur, counts = m.group(1), dict.fromkeys(ops, 0)
        continue
    if cur is None:
        continue
    m = re.search(r'/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if not m:
        continue
    op = m.group(1)
    base = op.split('.')[0].split('_')[0]
    if base in counts:
        counts[base] += 1
    if op.startswith('UTCHMMA.2CTA'):
        counts['UTCHMMA.2CTA'] += 1
    if base in ('ATOMS', 'ATOMG'):
        counts['ATOM'] += 1
if cur:
    rows.append((cur, counts))
dem = subprocess.run(['c++filt'], input='\n'.join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
print('# SASS opcode counts per kernel (`cuobjdump -sass dposer_b200/lib/libdposer_b200.so`, sm_100a)\n')
print('UTCHMMA = tcgen05.mma, .2CTA = cta_group::2; UTCBAR = tcgen05.commit; LDTM = tcgen05.ld; UTMALDG / UTMASTG = TMA load / store; '
      'SYNCS = mbarrier ops; UCGABAR = cluster barrier; ATOM = shared / global atomics with return, RED = reductions.\n')
print('| kernel | ' + ' | '.join(ops) + ' |')
print('|---|' + '---|' * len(ops))
tot = dict.fromkeys(ops, 0)
for (name, c), d in zip(rows, dem):
    short = re.sub(r'\(.*', '', d).replace('void ', '').replace('dpb::', '')
    print(f'| `{short[:70]}` | ' + ' | '.join(str(c[o]) for o in ops) + ' |')
    for o in ops:
        tot[o] += c[o]
print('| **total** | ' + ' | '.join(str(tot[o]) for o in ops) + ' |')
