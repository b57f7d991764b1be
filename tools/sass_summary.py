"""Per-kernel SASS evidence of the built library: counts of the Blackwell-native mnemonics (UTC*MMA = tcgen05.mma,
LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA, UBLKCP = bulk copy) and of the legacy tensor path (HMMA = mma.sync).
    python tools/sass_summary.py > profiles/rN_sass_summary.txt"""
import collections, os, re, subprocess, sys

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(HERE, 'ghn3_b200', 'libghn3_b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
kern, counts = None, collections.OrderedDict()
pat = re.compile(r'\b(UTC[A-Z]*MMA|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|UTCBAR|HMMA|IMMA|FFMA|MUFU|RED|ATOMG|SYNCS|LDGSTS)\b')
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        kern = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r'\(.*', '', kern)
        counts[kern] = collections.Counter()
        continue
    if kern:
        m = pat.search(line)
        if m:
            key = 'UTCMMA' if m.group(1).startswith('UTC') and m.group(1).endswith('MMA') else m.group(1)
            counts[kern][key] += 1
cols = ['UTCMMA', 'LDTM', 'STTM', 'UTMALDG', 'UBLKCP', 'HMMA', 'FFMA', 'MUFU', 'RED', 'SYNCS']
print('# cuobjdump -sass ghn3_b200/libghn3_b200.so (sm_100a): instruction counts per kernel')
print('# UTCMMA = tcgen05.mma (UTCHMMA ...), LDTM = tcgen05.ld, UTMALDG = TMA tensor load, HMMA = mma.sync')
print('%-86s' % 'kernel' + ''.join('%9s' % c for c in cols))
for k, c in counts.items():
    print('%-86s' % k[:85] + ''.join('%9d' % c.get(col, 0) for col in cols))
