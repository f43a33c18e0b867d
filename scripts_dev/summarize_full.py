"""ncu --set full report(s) -> markdown table of the metrics the roofline discussion uses.
usage: summarize_full.py <rep.ncu-rep> [...] """
import csv, io, subprocess, sys
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__shared_mem_per_block_dynamic']
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    for v in rows[2:]:
        print(f"\n**{v[h.index('Kernel Name')].split('(')[0]}** (`{rep.split('/')[-1]}`)\n\n| metric | value |\n|---|---|")
        for w in want:
            if w in h:
                i = h.index(w)
                print(f"| {w} | {v[i]} {units[i]} |")
