"""Attribute an ncu capture's per-SASS-instruction counters to CUDA source lines.

ncu's CSV export of the source page only carries the SASS view, so this joins it (by instruction order) with the
line table `nvdisasm -g` prints for the cubin of the SAME build (compile with -lineinfo).

usage: python tools/ncu_lines.py <report.ncu-rep> <kernel-name-substring> <lib.so> [launch-index] [top-n]
"""
import csv, io, re, subprocess, sys, tempfile, os, collections, glob


def sass_rows(rep, kernel, launch):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    blocks, cur = [], None
    for line in out.splitlines():
        if line.startswith('"Kernel Name"'):
            cur = {'name': line, 'lines': []}
            blocks.append(cur)
        elif cur is not None:
            cur['lines'].append(line)
    blocks = [b for b in blocks if kernel in b['name']]
    b = blocks[launch]
    rows = list(csv.reader(io.StringIO('\n'.join(b['lines']))))
    hdr = rows[0]
    iS, iW, iE = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
    return [(r[iS].strip(), int(r[iW] or 0), int(r[iE] or 0)) for r in rows[1:] if len(r) > iE]


def line_table(lib, kernel, n_expected=None):
    """Line table of the function whose name contains `kernel`; with several template instantiations, the one whose
    instruction count equals n_expected."""
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
    tables = {}
    for cub in glob.glob(os.path.join(tmp, '*.cubin')):
        txt = subprocess.run(['nvdisasm', '-g', cub], capture_output=True, text=True).stdout
        if kernel not in txt:
            continue
        in_fn, cur, res = False, ('?', 0, None), None
        for line in txt.splitlines():
            m = re.match(r'\s*\.section\s+\.text\.(\S+?),', line)
            if m:
                in_fn = kernel in m.group(1)
                if in_fn:
                    res = tables.setdefault(m.group(1), [])
                continue
            if not in_fn:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)),
                       (os.path.basename(m.group(3)), int(m.group(4))) if m.group(3) else None)
                continue
            m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
            if m:
                res.append((m.group(2).strip(), cur))
    if not tables:
        return []
    if n_expected is not None:
        for name, t in tables.items():
            if len(t) == n_expected:
                return t
    return max(tables.values(), key=len) if n_expected is None else min(tables.values(), key=lambda t: abs(len(t) - n_expected))


def main():
    rep, kernel, lib = sys.argv[1:4]
    launch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    rows = sass_rows(rep, kernel, launch)
    table = line_table(lib, kernel, len(rows))
    if len(rows) != len(table):
        print(f'WARNING: {len(rows)} profiled instructions vs {len(table)} in the cubin: different builds?')
    n = min(len(rows), len(table))
    mism = sum(1 for i in range(n) if rows[i][0].split()[0:1] != table[i][0].split()[0:1] and not rows[i][0].startswith('@'))
    agg = collections.defaultdict(lambda: [0, 0])
    tot_s = tot_e = 0
    for i in range(n):
        _, samples, execd = rows[i]
        f, l, outer = table[i][1]
        key = f'{f}:{l}' + (f'  <- {outer[0]}:{outer[1]}' if outer else '')
        agg[key][0] += samples; agg[key][1] += execd
        tot_s += samples; tot_e += execd
    print(f'{n} instructions, {tot_e} warp-instructions executed, {tot_s} samples, opcode mismatches: {mism}')
    print('--- by instructions executed')
    for k, (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
        print(f'{100*e/tot_e:6.2f}% inst {100*s/max(tot_s,1):6.2f}% samples  {k}')
    print('--- by stall samples')
    for k, (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
        print(f'{100*s/max(tot_s,1):6.2f}% samples {100*e/tot_e:6.2f}% inst  {k}')


if __name__ == '__main__':
    main()
