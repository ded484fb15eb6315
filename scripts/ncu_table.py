"""Condense an `ncu --set full --csv --page raw` log into one line per kernel launch with the
metrics DESIGN.md / profiles/ quote.  Usage: python scripts/ncu_table.py file.csv [more metrics]"""
import csv
import sys

KEYS = [
    ('dur_us', 'gpu__time_duration.sum', 1e-3),
    ('grid', 'launch__grid_size', 1),
    ('regs', 'launch__registers_per_thread', 1),
    ('tensor%', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 1),
    ('tensor_rt%', 'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 1),
    ('sm_thru%', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 1),
    ('dram_rd_MB', 'dram__bytes_read.sum', 1e-6),
    ('dram_wr_MB', 'dram__bytes_write.sum', 1e-6),
    ('dram%', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 1),
    ('l2->sm_MB', 'l1tex__m_xbar2l1tex_read_bytes.sum', 1e-6),
    ('l2_hit%', 'lts__t_sector_hit_rate.pct', 1),
    ('lsu_smem_wf%', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 1),
    ('ipc', 'sm__inst_executed.avg.per_cycle_elapsed', 1),
    ('warps/sm', 'sm__warps_active.avg.per_cycle_active', 1),
]


def main():
    path = sys.argv[1]
    extra = sys.argv[2:]
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
    col = {h: i for i, h in enumerate(hdr)}
    keys = [k for k in KEYS if k[1] in col] + [(m, m, 1) for m in extra if m in col]

    def num(r, name, scale):
        v = r[col[name]].replace(',', '')
        u = units[col[name]]
        try:
            f = float(v)
        except ValueError:
            return v
        if name.startswith('gpu__time_duration') and u in ('ns', 'nsecond'):
            f *= 1e-3
            return f
        if u == 'Mbyte': f *= 1e6
        if u == 'Kbyte': f *= 1e3
        if u == 'Gbyte': f *= 1e9
        return f * scale
    print('kernel'.ljust(34), ' '.join(k[0].rjust(11) for k in keys))
    for r in data:
        if len(r) < len(hdr):
            continue
        name = r[col['Kernel Name']].split('(')[0].replace('trb::<unnamed>::', '')[-34:]
        vals = []
        for k in keys:
            v = num(r, k[1], k[2])
            vals.append((f'{v:11.2f}' if isinstance(v, float) else str(v).rjust(11)))
        print(name.ljust(34), ' '.join(vals))


if __name__ == '__main__':
    main()
