"""Summarise an ncu report (.ncu-rep) into a small CSV for profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_axpy_sum.csv

One row per captured launch: duration, DRAM bytes read / written, DRAM throughput
as % of ncu's own peak, achieved occupancy, registers, executed instructions.
Runs in the CPU container (ncu -i needs no GPU)."""
import csv
import io
import subprocess
import sys

WANT = [
    ('Kernel Name', 'kernel'),
    ('Grid Size', 'grid'),
    ('Block Size', 'block'),
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'dram_read'),
    ('dram__bytes_write.sum', 'dram_write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct_of_ncu_peak'),
    ('lts__t_sector_hit_rate.pct', 'l2_hit_pct'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved_occupancy_pct'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__shared_mem_per_block_static', 'smem_static'),
    ('launch__shared_mem_per_block_dynamic', 'smem_dynamic'),
    ('smsp__inst_executed.sum', 'warp_insts'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_active_pct'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_pct'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem_bank_conflicts'),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(k), name) for k, name in WANT if k in hdr]
    with open(out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['%s[%s]' % (name, units[i]) if units[i] else name for i, name in cols])
        for r in rows[2:]:
            w.writerow([r[i][:160] for i, _ in cols])
    print(open(out).read())


if __name__ == '__main__':
    main()
