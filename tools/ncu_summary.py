"""Summarise an .ncu-rep (read here, without a GPU): duration, DRAM bytes / throughput, tensor pipe, occupancy, issue
slots, top stall reasons, pipe utilisation.  Usage: python tools/ncu_summary.py file.ncu-rep [more.ncu-rep ...]"""
import csv, io, re, subprocess, sys

KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_subunit_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_tensor.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.per_cycle_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.sum',
    'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_lsu.sum',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg',
]


def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    return hdr, units, vals


for path in sys.argv[1:]:
    hdr, units, vals = raw(path)
    for v in vals:
        name = v[hdr.index('Kernel Name')]
        print(f'## {path}: {name[:100]}')
        d = {h: (x, u) for h, u, x in zip(hdr, units, v)}
        for k in KEYS:
            if k in d:
                print(f'  {k:75s} {d[k][0]:>16s} {d[k][1]}')
        pipes = [(h, d[h][0]) for h in hdr if h.startswith('sm__inst_executed_pipe_') and h.endswith('.avg.pct_of_peak_sustained_active')]
        for h, x in sorted(pipes, key=lambda t: -float(t[1].replace(',', '') or 0))[:6]:
            print(f'  {h:75s} {x:>16s} %')
        stalls = [(h, d[h][0]) for h in hdr if re.match(r'smsp__average_warps?_issue_stalled_.*_per_issue_active\.ratio', h) or
                  re.match(r'smsp__average_warp_latency_issue_stalled_.*\.ratio', h)]
        st = []
        for h, x in stalls:
            try:
                st.append((float(x.replace(',', '')), h))
            except ValueError:
                pass
        for x, h in sorted(st, reverse=True)[:7]:
            print(f'  stall {h.split("issue_stalled_")[1].split("_per")[0].replace(".ratio", ""):30s} {x:8.2f}')
