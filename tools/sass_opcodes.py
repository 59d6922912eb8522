"""Per-kernel counts of the Blackwell-specific SASS opcodes in the in-tree library (cuobjdump -sass):
UTCHMMA (tcgen05.mma kind::f16), LDTM / STTM (tcgen05.ld / st), UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit),
UTCATOM* (TMEM alloc), SYNCS (mbarrier), MUFU, HFMA2 / HMUL2 / HADD2, DADD, FFMA, RED/ATOM.

    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, '2024-hl-spi3s-sunerf_b200', 'lib', 'libsunerf_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(['cu++filt', n], capture_output=True, text=True).stdout.strip() or n
ops = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UBLKCP', 'UTCBAR', 'UTCATOM', 'SYNCS', 'MUFU', 'HFMA2', 'HMUL2', 'HADD2', 'HMNMX2', 'DADD',
       'FFMA', 'RED', 'ATOM', 'LDS', 'STS', 'LDG', 'STG', 'SHFL']
cur, counts, total = None, collections.OrderedDict(), {}
for line in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); total[cur] = 0
        continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and cur:
        total[cur] += 1
        op = m.group(1)
        for o in ops:
            if op == o or op.startswith(o + '.') or (o == 'UTCATOM' and op.startswith('UTCATOM')):
                counts[cur][o] += 1
print(f'# {os.path.relpath(lib, ROOT)}: SASS opcode counts per kernel (sm_100a), cuobjdump -sass')
print('# ' + ' '.join(f'{o:>7s}' for o in ['instrs'] + ops) + '  kernel')
for k, c in counts.items():
    name = demangle(k)
    name = re.sub(r'\((?!bool|int)[^)]*\)$', '', name).replace('snf::bf::', '').replace('snf::', '').replace('(anonymous namespace)::', '')
    print('  ' + ' '.join(f'{v:7d}' for v in [total[k]] + [c[o] for o in ops]) + '  ' + name)
