"""Per-kernel SASS instruction census of libssdnerf_b200.so (cuobjdump -sass): the mnemonics that prove which hardware path a kernel
uses -- UTCHMMA (tcgen05.mma), UTMALDG / UTMASTG (TMA), LDTM / STTM (TMEM load / store), HMMA (legacy mma.sync), MUFU, FFMA ...
usage: python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'ssdnerf_b200', 'libssdnerf_b200.so')
KEYS = ['UTCHMMA', 'UTCQMMA', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'LDTM', 'STTM', 'UTCBAR', 'SYNCS', 'HMMA', 'MUFU', 'FFMA', 'LDG', 'STG', 'LDS', 'STS', 'ATOM', 'RED', 'BAR']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    kern, counts, order = None, collections.defaultdict(collections.Counter), []
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            kern = subprocess.run(['c++filt', m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            kern = re.sub(r'\(.*', '', kern).replace('ssdnerf::', '')
            order.append(kern)
            continue
        m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m and kern:
            op = m.group(1).split('.')[0]
            counts[kern]['total'] += 1
            if op in KEYS:
                counts[kern][op] += 1
    used = [k for k in KEYS if any(counts[n][k] for n in order)]
    print(f'# {os.path.relpath(LIB, ROOT)}  (cuobjdump -sass, sm_100a); instruction counts per kernel; blank = 0')
    print('%-72s %7s ' % ('kernel', 'total') + ' '.join('%7s' % k for k in used))
    for n in sorted(set(order)):
        print('%-72s %7d ' % (n[:72], counts[n]['total']) + ' '.join('%7s' % (counts[n][k] or '') for k in used))


if __name__ == '__main__':
    main()
