"""Summarise an .ncu-rep (raw page) per kernel: duration, DRAM bytes, throughputs, stalls."""
import csv, subprocess, sys, io
rep = sys.argv[1]
every = int(sys.argv[2]) if len(sys.argv) > 2 else 1
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {n: i for i, n in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum']
stalls = [h for h in hdr if 'issue_stalled' in h and 'ratio' in h and 'not_issued' not in h]
for k, r in enumerate(data):
    if k % every:
        continue
    print('==', r[idx['Kernel Name']][:80])
    for w in want:
        if w in idx:
            print(f'   {w} [{units[idx[w]]}] = {r[idx[w]]}')
    st = sorted([(float(r[idx[n]] or 0), n) for n in stalls], reverse=True)[:6]
    print('   stalls/issue: ' + ', '.join('%s %.2f' % (n.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v) for v, n in st))
