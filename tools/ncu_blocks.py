"""Summarise an `ncu --page raw --csv` + `--page source --csv` export pair: headline metrics and an instruction /
stall-sample histogram over blocks of SASS instructions (usage: ncu_blocks.py raw.csv src.csv [block])."""
import csv, collections, sys
raw, src = sys.argv[1], sys.argv[2]
B = int(sys.argv[3]) if len(sys.argv) > 3 else 100
rows = list(csv.reader(open(raw)))
hdr = rows[0]; vals = rows[2]
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum']
for i, h in enumerate(hdr):
    try:
        big = 'issue_stalled' in h and 'per_issue_active' in h and float(vals[i] or 0) > 0.3
    except ValueError:
        big = False
    if h in want or big:
        print(h, vals[i])
rows = list(csv.reader(open(src)))
hdr = rows[1]; data = rows[2:]
ia = hdr.index('Instructions Executed'); isrc = hdr.index('Source'); isamp = hdr.index('# Samples')
tot = sum(int(r[ia]) for r in data); ts = sum(int(r[isamp]) for r in data)
print('total inst', tot, 'n sass', len(data), 'samples', ts)
for b in range(0, len(data), B):
    blk = data[b:b + B]
    n = sum(int(r[ia]) for r in blk); sm = sum(int(r[isamp]) for r in blk)
    if n / tot < 0.01 and sm / ts < 0.01:
        continue
    ops = collections.Counter()
    for r in blk:
        t = r[isrc].split(); op = t[1] if t[0].startswith('@') else t[0]
        ops[op.split('.')[0]] += int(r[ia])
    top = ', '.join('%s %.0f%%' % (k, 100 * v / max(n, 1)) for k, v in ops.most_common(5))
    print('%5d  inst %5.1f%%  samples %5.1f%%  %s' % (b, 100 * n / tot, 100 * sm / ts, top))
