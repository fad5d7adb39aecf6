"""Key raw metrics of an ncu report: python tools/ncu_key.py rep.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'lts__t_bytes.sum', 'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed']
for vals in rows[2:]:
    print("kernel:", vals[hdr.index('Kernel Name')][:60] if 'Kernel Name' in hdr else '')
    for i, h in enumerate(hdr):
        if h in want:
            print("  ", h, units[i], vals[i])
