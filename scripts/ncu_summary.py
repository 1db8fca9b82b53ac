"""`ncu -i X.ncu-rep --page raw --csv` -> the one-line-per-launch summary kept under profiles/:
    python scripts/ncu_summary.py gpurun_out/r02_targets.ncu-rep > profiles/r02_targets_ncu_full.csv
"""
import csv
import io
import subprocess
import sys

COLS = [
    ('duration[us]', 'gpu__time_duration.sum'),
    ('dram_read[Gbyte]', 'dram__bytes_read.sum'),
    ('dram_write[Gbyte]', 'dram__bytes_write.sum'),
    ('dram_pct_of_ncu_peak[%]', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
    ('l2_hit_pct[%]', 'lts__t_sector_hit_rate.pct'),
    ('achieved_occupancy_pct[%]', 'sm__warps_active.avg.pct_of_peak_sustained_active'),
    ('regs[register/thread]', 'launch__registers_per_thread'),
    ('smem_dynamic[Kbyte/block]', 'launch__shared_mem_per_block_dynamic'),
    ('warp_insts[inst]', 'smsp__inst_executed.sum'),
    ('issue_active_pct[%]', 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
    ('smem_bank_conflicts', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    w = csv.writer(sys.stdout)
    w.writerow(['kernel', 'grid', 'block'] + [c for c, _ in COLS])
    for r in rows[2:]:
        def val(metric):
            i = idx.get(metric)
            if i is None:
                return ''
            v = r[i]
            u = units[i]
            try:
                f = float(v)
            except ValueError:
                return v
            if metric.startswith('dram__bytes'):          # normalise to Gbyte
                f *= {'byte': 1e-9, 'Kbyte': 1e-6, 'Mbyte': 1e-3, 'Gbyte': 1.0, 'Tbyte': 1e3}.get(u, 1.0)
            if metric == 'gpu__time_duration.sum':
                f *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(u, 1.0)
            return '%.6g' % f
        name = r[idx['Kernel Name']]
        w.writerow([name[:160], r[idx['Grid Size']], r[idx['Block Size']]] + [val(m) for _, m in COLS])


if __name__ == '__main__':
    main()
