"""Summarise an `ncu --set full` report into the few numbers DESIGN.md / bench.py quote.

usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.txt
Per profiled launch: duration, DRAM bytes (read+write = `roofline.traffic`), DRAM / L2 / tensor-pipe
utilisation as ncu reports them, registers, grid.  Reading the report needs no GPU.
"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__bytes_read.sum.per_second", "dram read rate"),
    ("dram__bytes_write.sum.per_second", "dram write rate"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of ncu peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("sm__cycles_elapsed.avg", "SM cycles"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}: {len(rows) - 2} profiled launch(es)  (ncu --set full --clock-control none)")
    for r in rows[2:]:
        print(f"\nkernel: {r[idx['Kernel Name']][:110]}")
        rd = wr = None
        for key, label in WANT:
            if key in idx:
                print(f"  {label:26s} {r[idx[key]]:>16s} {units[idx[key]]}")
                if key == "dram__bytes_read.sum":
                    rd = (float(r[idx[key]]), units[idx[key]])
                if key == "dram__bytes_write.sum":
                    wr = (float(r[idx[key]]), units[idx[key]])
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        if rd and wr:
            print(f"  {'traffic (read+write)':26s} {(rd[0] * scale[rd[1]] + wr[0] * scale[wr[1]]) / 1e6:16.3f} MB")


if __name__ == "__main__":
    main(sys.argv[1])
