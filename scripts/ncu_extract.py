"""Reduce an `ncu --set full` report (raw page CSV) to the metrics the profiles/ summaries quote.

    ncu -i rep.ncu-rep --page raw --csv > raw.csv ; python scripts/ncu_extract.py raw.csv out.csv
"""
import csv
import re
import sys

KEEP = re.compile(
    r"^(ID|Kernel Name|Block Size|Grid Size|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum(\.per_second)?|"
    r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|dram__cycles_active\.avg\.pct_of_peak_sustained_elapsed|"
    r"lts__throughput\.avg\.pct_of_peak_sustained_elapsed|lts__t_bytes\.sum|"
    r".*sm__pipe_tensor_cycles_active.*|.*pipe_tensor_subpipe_hmma_cycles_active.*|"
    r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
    r"launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_dynamic|occupancy_limit_.*)|"
    r"sm__cycles_elapsed\.avg(\.per_second)?|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum(\.pct_of_peak_sustained_elapsed)?|"
    r"smsp__inst_executed\.sum|sm__inst_executed_pipe_(fma|alu|lsu|uniform)\.sum)$")


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [i for i, h in enumerate(hdr) if KEEP.match(h)]
    with open(sys.argv[2], "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{j}" for j in range(len(data))])
        for i in cols:
            w.writerow([hdr[i], units[i]] + [r[i] for r in data])


if __name__ == "__main__":
    main()
