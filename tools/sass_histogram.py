"""SASS instruction histogram per kernel of libp3d.so (cuobjdump -sass), written as a markdown table.

usage: python tools/sass_histogram.py [pix2pix3d_b200/libp3d.so] > profiles/rNN_sass_histogram.md

The columns are the mnemonics that show which hardware path a kernel uses: UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM
(tcgen05.ld / st), UTMALDG (TMA tensor loads), UBLKCP (bulk copies), SYNCS (mbarrier), FFMA2 / FADD2 / FMUL2 (packed fp32),
HFMA2, MUFU, LDG / STG / LDS / STS, ATOMG / RED, BAR.
"""
import collections
import re
import subprocess
import sys

COLS = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UBLKCP', 'SYNCS', 'FFMA2', 'FADD2', 'FMUL2', 'FFMA', 'HFMA2', 'HADD2', 'MUFU',
        'LDG', 'STG', 'LDS', 'STS', 'ATOMG', 'RED', 'ATOMS', 'BAR', 'SHFL']


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else 'pix2pix3d_b200/libp3d.so'
    txt = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur['_total'] += 1
    names = demangle(list(kernels))
    print('# SASS instruction histogram per kernel (`cuobjdump -sass %s`, sm_100a)\n' % path)
    print('Static counts per kernel (all template instances listed). `total` = instructions in the kernel body.\n')
    used = [c for c in COLS if any(k[c] for k in kernels.values())]
    print('| kernel | total | ' + ' | '.join(used) + ' |')
    print('|---|---:|' + '---:|' * len(used))
    def short(n):
        d = names.get(n, n)
        d = re.sub(r'^void ', '', d)
        d = re.sub(r'\((.|\n)*$', '', d)
        return d.replace('p3d::', '').replace('(anonymous namespace)::', '')
    for n, c in sorted(kernels.items(), key=lambda kv: short(kv[0])):
        print('| `%s` | %d | ' % (short(n), c['_total']) + ' | '.join(str(c[x]) if c[x] else '' for x in used) + ' |')


if __name__ == '__main__':
    main()
