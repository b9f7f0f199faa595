"""Turns an `ncu --set full` report into the per-kernel table and the traffic file bench.py reads.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv      (done by this script)
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r2_ncu_summary.md profiles/r2_ncu_traffic.json 'title' batch=8192,n_nodes=1152,n_caps=43

One row per kernel (template arguments kept), means over its launches.
"""
import collections
import csv
import json
import re
import subprocess
import sys

COLS = {
    'time_ms': 'gpu__time_duration.sum',
    'regs': 'launch__registers_per_thread',
    'grid': 'launch__grid_size',
    'block': 'launch__block_size',
    'issue': 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'tensor': 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'fma': 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'smem': 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
    'dram_rd': 'dram__bytes_read.sum',
    'dram_wr': 'dram__bytes_write.sum',
    'dram_pct': 'dram__bytes_read.sum.pct_of_peak_sustained_elapsed',
    'dram_pct_w': 'dram__bytes_write.sum.pct_of_peak_sustained_elapsed',
    'inst': 'smsp__inst_executed.sum',
}
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12,
         'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3, 'usecond': 1e-3, 'msecond': 1.0, 'nsecond': 1e-6, 'second': 1e3}


def short(name):
    name = re.sub(r'\(int\)', '', name)
    name = re.sub(r'\(bool\)', '', name)
    m = re.search(r'(k_[A-Za-z0-9_]+)(<[^>]*>)?', name)
    return (m.group(1) + (m.group(2) or '')).replace(' ', '') if m else name[:40]


def main():
    rep, out_md, out_json = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else ''
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {k: hdr.index(v) for k, v in COLS.items()}
    iname = hdr.index('Kernel Name')
    agg = collections.OrderedDict()
    for r in data:
        k = short(r[iname])
        a = agg.setdefault(k, collections.defaultdict(float))
        a['n'] += 1
        for key, i in ix.items():
            v = float(r[i].replace(',', '')) if r[i] not in ('', 'n/a') else 0.0
            a[key] += v * SCALE.get(units[i], 1.0)
    with open(out_md, 'w') as f:
        f.write('# %s\n# ncu --set full --clock-control none; one row per kernel, means over its launches (tools/ncu_summary.py)\n\n' % title)
        f.write('| kernel | launches | time ms | regs | grid x block | issue active % | tensor active % | FMA pipe % | smem pipe % | '
                'DRAM read GB | DRAM write GB | DRAM % of peak | warp instr (M) |\n|' + '---|' * 13 + '\n')
        for k, a in agg.items():
            n = a['n']
            f.write('| %s | %d | %.3f | %d | %d x %d | %.1f | %.1f | %.1f | %.1f | %.3f | %.3f | %.1f | %.1f |\n' % (
                k, n, a['time_ms'] / n, a['regs'] / n, a['grid'] / n, a['block'] / n, a['issue'] / n, a['tensor'] / n,
                a['fma'] / n, a['smem'] / n, a['dram_rd'] / n / 1e9, a['dram_wr'] / n / 1e9, (a['dram_pct'] + a['dram_pct_w']) / n, a['inst'] / n / 1e6))
    kern = {k: {'dram_bytes_per_launch': (a['dram_rd'] + a['dram_wr']) / a['n'], 'time_ms_under_ncu': a['time_ms'] / a['n'],
                'launches': int(a['n'])} for k, a in agg.items()}
    # bench.py's kernel classes (KCLASS indices) -> measured DRAM bytes per launch, averaged over the class's launches
    cls_of = [(1, 'k_pass_tc<0'), (2, 'k_pass_tc<1'), (3, 'k_pass_tc<2'), (10, 'k_sweep_fused'), (6, 'k_grad'), (11, 'k_c1_')]
    by_cls = {}
    for cid, prefix in cls_of:
        tot = sum((a['dram_rd'] + a['dram_wr']) for k, a in agg.items() if k.startswith(prefix))
        cnt = sum(a['n'] for k, a in agg.items() if k.startswith(prefix))
        if cnt:
            by_cls[str(cid)] = tot / cnt
    step_total = sum(a['dram_rd'] + a['dram_wr'] for a in agg.values())
    dims = {}
    for kv in (sys.argv[5].split(',') if len(sys.argv) > 5 else []):      # e.g. batch=8192,n_nodes=1152,n_caps=43
        k, v = kv.split('=')
        dims[k] = int(v)
    out = {'source': '%s (ncu --set full --clock-control none), via tools/ncu_summary.py' % out_md, 'title': title}
    out.update(dims)
    out.update({'dram_bytes_all_captured_launches': step_total, 'dram_bytes_per_launch_by_class': by_cls, 'kernels': kern})
    json.dump(out, open(out_json, 'w'), indent=1)
    print(open(out_md).read())


if __name__ == '__main__':
    main()
