"""Key metrics of one kernel launch in an ncu report as a markdown table: python tools/ncu_summary.py <report.ncu-rep> [launch-index]"""
import csv, subprocess, sys
rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2 + idx]
get = dict(zip(hdr, zip(vals, units)))
WANT = [
    ('Kernel Name', 'kernel'), ('Grid Size', 'grid'), ('Block Size', 'block'), ('launch__registers_per_thread', 'registers / thread'),
    ('launch__occupancy_limit_registers', 'CTAs / SM (register limit)'), ('launch__occupancy_limit_shared_mem', 'CTAs / SM (shared-memory limit)'),
    ('gpu__time_duration.sum', 'duration'), ('smsp__inst_executed.sum', 'warp instructions executed'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
    ('smsp__warps_eligible.avg.per_cycle_active', 'eligible warps / scheduler / cycle'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
    ('sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'ALU pipe %'),
    ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'FMA pipe %'),
    ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'FMA pipe cycles %'),
    ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'XU pipe %'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_active', 'L1/TEX throughput %'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput %'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput %'),
    ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM written'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'shared-memory bank conflicts'),
]
print('| metric | value |\n|---|---|')
for key, label in WANT:
    if key in get:
        v, u = get[key]
        print(f'| {label} | {v} {u} |'.replace('  |', ' |'))
stalls = sorted(((float(v[0]), k) for k, v in get.items() if 'issue_stalled' in k and k.endswith('_per_issue_active.ratio') and v[0]), reverse=True)
print('\nWarp stall reasons (warps stalled per issued instruction): ' + ', '.join(
    f"{k.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')} {v:.2f}" for v, k in stalls[:8] if 'selected' not in k.split('issue_stalled_')[1][:8]))
