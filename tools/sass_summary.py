"""Per-kernel SASS opcode summary of libbear_b200.so (cuobjdump -sass): the instructions that prove which hardware
paths a kernel uses -- TMA bulk copies (UBLKCP), tcgen05 (UTCIMMA / UTCBAR / LDTM / UTCATOMSWS alloc), mbarriers (SYNCS),
FP64 tensor cores (DMMA), and the ones that should be absent (MATCH, shared-memory atomics ATOMS).
    python tools/sass_summary.py > profiles/r2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'bear_b200', 'libbear_b200.so')
KEYS = ['UBLKCP', 'UTCIMMA', 'UTCBAR', 'LDTM', 'UTCATOMSWS', 'SYNCS', 'DMMA', 'DFMA', 'MATCH', 'ATOMS', 'ATOMG', 'RED', 'LDL', 'STL']
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
name, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r'\(anonymous namespace\)::', '', name)
        name = re.sub(r'\(.*', '', name)
        counts[name] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and name:
        counts[name]['total'] += 1
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + '.') or (k == 'RED' and op == 'RED'):
                counts[name][k] += 1
print('# static SASS opcode counts per kernel of %s (sm_100a); blank = 0' % os.path.basename(lib))
print('%-78s %7s ' % ('kernel', 'instr') + ' '.join('%7s' % k[:7] for k in KEYS))
for n, c in counts.items():
    print('%-78s %7d ' % (n[:78], c['total']) + ' '.join('%7s' % (c[k] or '') for k in KEYS))
