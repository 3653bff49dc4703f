#!/usr/bin/env python
"""Turn ncu artefacts brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv  > profiles/r01_launches.txt
    python tools/summarize_ncu.py kernel   gpurun_out/prof.ncu-rep     > profiles/r01_ensemble_tc.txt
"""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith('==')))
    hdr = rows[0]
    ik, iv, im = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
    tot = collections.OrderedDict()
    n = collections.Counter()
    for r in rows[1:]:
        if len(r) <= iv or r[im] != 'gpu__time_duration.sum':
            continue
        name = r[ik].split('(')[0]
        tot[name] = tot.get(name, 0.0) + float(r[iv].replace(',', ''))
        n[name] += 1
    total = sum(tot.values())
    print('# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)')
    print('%-60s %8s %14s %8s' % ('kernel', 'launches', 'total ns', 'share'))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print('%-60s %8d %14.0f %7.2f%%' % (k[-60:], n[k], v, 100 * v / total))


KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.per_cycle_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed']


def kernel(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print('kernel: %s   grid %s block %s' % (d.get('Kernel Name', '?'), d.get('Grid Size'), d.get('Block Size')))
        for k in KEYS:
            if k in d:
                print('  %-82s %18s %s' % (k, d[k], units[hdr.index(k)]))
        for k in hdr:
            if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k:
                v = float(d[k] or 0)
                if v > 0.15:
                    print('  stall %-40s %.2f warps per issue' % (
                        k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))


def traffic(path):
    """DRAM bytes (read + write) of the captured launch as a JSON object, for bench.py's roofline.traffic."""
    import json
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    d = dict(zip(hdr, rows[2]))
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}

    def val(k):
        return float(d[k].replace(',', '')) * scale[units[hdr.index(k)]]
    print(json.dumps({'kernel': d.get('Kernel Name'), 'dram_bytes_read': val('dram__bytes_read.sum'),
                      'dram_bytes_write': val('dram__bytes_write.sum'),
                      'gpu_time_ms': float(d['gpu__time_duration.sum'].replace(',', '')) *
                      {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}[units[hdr.index('gpu__time_duration.sum')]],
                      'source': 'ncu --set full --clock-control none, one launch of the bench workload'}))


if __name__ == '__main__':
    {'launches': launches, 'kernel': kernel, 'traffic': traffic}[sys.argv[1]](sys.argv[2])
