"""Summarise an .ncu-rep (raw page) : one block of key metrics per kernel launch."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_active.avg', 'sm__cycles_elapsed.max',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_fmaheavy.sum', 'sm__inst_executed_pipe_fmalite.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_bytes.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active']
stalls = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    print('====', r[idx['Kernel Name']][:70])
    for w in want:
        if w in idx:
            print(f'  {w:72s} {r[idx[w]]:>16s} {units[idx[w]]}')
    st = sorted(((float(r[idx[h]].replace(",", "")), h) for h in stalls if r[idx[h]] not in ('', 'n/a')), reverse=True)[:7]
    for v, h in st:
        print(f'  stall {h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]:40s} {v:8.3f}')
