"""Extract per-launch DRAM traffic of a kernel from `ncu --set full` captures into
profiles/ncu_traffic.json (read by bench.py for `roofline.traffic`).

usage: python tools/ncu_traffic.py <kernel-name> <report.ncu-rep> <algorithmic bytes of launch 1> [<launch 2> ...]
The algorithmic byte counts are those of the PROFILED launches, in capture order.
"""
import csv
import json
import subprocess
import sys
from pathlib import Path

SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    kernel, rep = sys.argv[1], sys.argv[2]
    algo = [float(x) for x in sys.argv[3:]]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r, a in zip(rows[2:], algo):
        rd = float(r[idx["dram__bytes_read.sum"]]) * SCALE[units[idx["dram__bytes_read.sum"]]]
        wr = float(r[idx["dram__bytes_write.sum"]]) * SCALE[units[idx["dram__bytes_write.sum"]]]
        us = float(r[idx["gpu__time_duration.sum"]])
        launches.append({"dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr, "algorithmic_bytes": a,
                         "duration": us, "duration_unit": units[idx["gpu__time_duration.sum"]]})
    p = Path(__file__).resolve().parent.parent / "profiles" / "ncu_traffic.json"
    d = json.loads(p.read_text()) if p.exists() else {}
    d[kernel] = {"source": f"ncu --set full --clock-control none, {Path(rep).name}", "launches": launches}
    p.write_text(json.dumps(d, indent=1))
    print(json.dumps(d[kernel], indent=1))


if __name__ == "__main__":
    main()
