"""Per-kernel stall picture from an `ncu --set full --import-source on` report: the SASS lines that collect the most
stall samples, with their dominant stall reasons, plus the instruction mix of the hot loop.

    python tools/ncu_stalls.py gpurun_out/r2_full.ncu-rep k_sweep_fused [top_n]
"""
import csv
import collections
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + pat, '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'hdr': None, 'rows': []}
            blocks.append(cur)
        elif cur is not None and r and r[0] == 'Address':
            cur['hdr'] = r
        elif cur is not None and cur['hdr'] and len(r) == len(cur['hdr']):
            cur['rows'].append(r)
    seen = set()
    for b in blocks:
        if b['name'] in seen:
            continue
        seen.add(b['name'])
        h = b['hdr']
        i_src, i_samp, i_exec = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
        stall_cols = [(i, c[6:]) for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
        total = sum(int(r[i_samp] or 0) for r in b['rows'])
        tot_by = collections.Counter()
        for r in b['rows']:
            for i, c in stall_cols:
                tot_by[c] += int(r[i] or 0)
        print('## %s\nsamples %d; by reason: %s' % (b['name'][:90], total, ', '.join('%s %.1f%%' % (k, 100.0 * v / max(total, 1)) for k, v in tot_by.most_common(9))))
        mx = max(int(r[i_exec] or 0) for r in b['rows'])
        mix = collections.Counter()
        for r in b['rows']:
            e = int(r[i_exec] or 0)
            if e >= mx * 0.4:
                op = r[i_src].split()[0] if not r[i_src].strip().startswith('@') else r[i_src].split()[1]
                mix[op.split('.')[0]] += e
        tot_hot = sum(mix.values())
        print('hot-loop instruction mix (lines executed >= 40%% of the max): ' + ', '.join('%s %.1f%%' % (k, 100.0 * v / tot_hot) for k, v in mix.most_common(14)))
        print('| share of samples | executed | SASS | top stall reasons |\n|---|---|---|---|')
        for r in sorted(b['rows'], key=lambda r: -int(r[i_samp] or 0))[:top]:
            s = int(r[i_samp] or 0)
            rs = sorted(((int(r[i] or 0), c) for i, c in stall_cols), reverse=True)[:2]
            print('| %.1f %% | %s | `%s` | %s |' % (100.0 * s / max(total, 1), r[i_exec], ' '.join(r[i_src].split())[:70],
                                                   ', '.join('%s %d%%' % (c, 100 * v / max(s, 1)) for v, c in rs)))
        print()


if __name__ == '__main__':
    main()
