"""Text summary of an .ncu-rep (one block per profiled launch) for profiles/: usage ncu_summary.py rep [rep ...] > out.txt"""
import csv, subprocess, sys, io
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct', 'dram__throughput.avg.pct_of_peak_sustained_elapsed']
print("# ncu --set full --clock-control none --import-source on (one block per profiled launch); workload: bench C2, 16 Mi-sample chunk")
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("---")
        for i, h in enumerate(hdr):
            stall = 'issue_stalled' in h and 'per_issue_active' in h
            try:
                keep = h == 'Kernel Name' or h in WANT or (stall and float(r[i] or 0) > 0.05)
            except ValueError:
                keep = False
            if keep:
                print("%-92s %s %s" % (h, r[i], units[i]))
