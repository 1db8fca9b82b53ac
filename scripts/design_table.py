"""Renders the measurement table of DESIGN.md section 7 from committed bench lines (profiles/r02_bench_*.json):
    python scripts/design_table.py profiles/r02_bench_n1.json [profiles/r02_bench_n2.json ...]
"""
import json
import sys


def load(path):
    with open(path) as f:
        txt = f.read()
    for line in txt.splitlines():
        line = line.strip()
        if line.startswith('{'):
            return json.loads(line)
    raise SystemExit('no JSON line in %s' % path)


def main():
    lines = [load(p) for p in sys.argv[1:]]
    n1 = next((d for d in lines if d.get('n_gpus') == 1), None)
    if n1:
        print('| config | ms | GB/s | % of measured peak | parity check (in-bench) | reference GPU kernels, same run | speed-up |')
        print('|---|---|---|---|---|---|---|')
        for e in n1.get('configs', []):
            if 'ms' not in e:
                print('| %s | host %.1f us / call (sum), %.1f us (add); with an event pair per call %.1f us; in a CUDA graph %s us | | | %s | %s | |' % (
                    e['name'], e['host_us_per_call_sum'], e['host_us_per_call_add'], e['gpu_us_per_call_sum'],
                    ('%.1f' % e['graph_replay_us_sum']) if e.get('graph_replay_us_sum') is not None else '-',
                    e['check'], e.get('reference_documented', '')))
                continue
            r = e.get('ref_gpu') or {}
            ref = ('%.3f ms, %.0f GB/s' % (r['ms'], r['gbs'])) if 'ms' in r else (r.get('skipped') or r.get('error') or '')
            name = e['name']
            if e.get('graph_replay_us') is not None:
                name += ' -- host-bound per call; the kernel alone, replayed in a CUDA graph: %.1f us = %.0f GB/s (%.0f %%)' % (
                    e['graph_replay_us'], e['graph_replay_gbs'], 100 * e['graph_replay_frac'])
            print('| %s | %.3f | %.0f | %.1f | %s | %s | %s |' % (
                name, e['ms'], e['gbs'], 100 * e['frac'], e['check'], ref,
                ('%.2fx' % e['speedup_vs_ref_gpu']) if 'speedup_vs_ref_gpu' in e else ''))
        c1 = n1.get('c1_cpu')
        if c1:
            print()
            print('C1 on the host (NumPy %s, 1 of %d cores): x*2+1 best %.2f ms / median %.2f ms (%.1f GB/s); '
                  'x.sum(axis=1) best %.2f ms / median %.2f ms (%.1f GB/s).' % (
                      c1['numpy'], c1['host_cpu_count'], c1['x*2+1']['best_ms'], c1['x*2+1']['median_ms'],
                      c1['x*2+1']['gbs_best'], c1['x.sum(axis=1)']['best_ms'], c1['x.sum(axis=1)']['median_ms'],
                      c1['x.sum(axis=1)']['gbs_best']))
    print()
    print('| N | headline GB/s (weak) | ms/step | per-GPU % of peak | e2e GB/s | C5 sum ms (GB/s aggregate) | C5 var ms (GB/s aggregate) | combine |')
    print('|---|---|---|---|---|---|---|---|')
    base = {}
    for d in sorted(lines, key=lambda d: d['n_gpus']):
        c = d.get('c5', {})
        n = d['n_gpus']
        if n == 1 and c:
            base = {k: c[k]['ms'] for k in ('sum', 'var')}
        eff = lambda k: (' eff %.3f' % (base[k] / (n * c[k]['ms']))) if base and c else ''
        print('| %d | %.0f | %.4f | %.1f | %.1f | %s | %s | %s |' % (
            n, d['value'], d['ms_per_step'], d['pct_of_peak']['measured'], d['e2e']['value'],
            ('%.3f (%.0f)%s' % (c['sum']['ms'], c['sum']['gbs_aggregate'], eff('sum'))) if c else '',
            ('%.3f (%.0f)%s' % (c['var']['ms'], c['var']['gbs_aggregate'], eff('var'))) if c else '',
            (c.get('combine', '')[:40]) if c else ''))


if __name__ == '__main__':
    main()
