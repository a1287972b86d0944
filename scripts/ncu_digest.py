"""Digest of an `ncu --set full` report: per captured launch the duration, DRAM traffic, pipe utilisation and the
warp-stall breakdown.  usage: ncu -i X.ncu-rep --page raw --csv | python scripts/ncu_digest.py"""
import csv, re, sys
rows = list(csv.reader(sys.stdin))
hdr, units, data = rows[0], rows[1], rows[2:]
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_active.avg', 'sm__cycles_elapsed.max',
        'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__sass_l1tex_data_pipe_lsu_wavefronts_mem_shared_op_ldgsts.sum', 'smsp__inst_executed.sum',
        'smsp__warps_active.avg.per_cycle_active', 'smsp__issue_active.avg.per_cycle_active']
for n, r in enumerate(data):
    print("launch %d: %s" % (n, r[hdr.index('Kernel Name')]))
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("  %-82s %14s %s" % (k, r[i], units[i]))
    st = []
    for i, h in enumerate(hdr):
        m = re.match(r'smsp__average_warps_issue_stalled_(.*)_per_issue_active\.ratio$', h)
        if m:
            st.append((float(r[i]), m.group(1)))
    tot = sum(v for v, _ in st) or 1.0
    print("  warp stalls per issue (share): " + ", ".join("%s %.2f (%.0f%%)" % (k, v, 100 * v / tot) for v, k in sorted(st, reverse=True)[:8]))
