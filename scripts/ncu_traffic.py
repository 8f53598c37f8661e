"""profiles/ncu_traffic.json from `ncu --set full` captures: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch of
every kernel in the given .ncu-rep files, averaged over the captured launches, keyed "<kernel>@<cfg>" (bench.py's roofline.traffic).
usage: ncu_traffic.py cfg4:gpurun_out/x/prof_cfg4.ncu-rep [cfg5:...raw.csv]   (.ncu-rep: runs `ncu -i ... --page raw --csv`; or that csv itself)"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def units(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


def main():
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    table = {}
    meta = {}
    for arg in sys.argv[1:]:
        cfg, rep = arg.split(":", 1)
        raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, unit = rows[0], rows[1]
        ix = {k: i for i, k in enumerate(hdr)}
        acc = {}
        for r in rows[2:]:
            name = re.sub(r"^(void )?(gnf::)?", "", r[ix["Kernel Name"]]).split("(")[0].split("<")[0]
            rd = units(r[ix["dram__bytes_read.sum"]], unit[ix["dram__bytes_read.sum"]])
            wr = units(r[ix["dram__bytes_write.sum"]], unit[ix["dram__bytes_write.sum"]])
            us = float(r[ix["gpu__time_duration.sum"]].replace(",", ""))
            tp = r[ix["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]] if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed" in ix else None
            acc.setdefault(name, []).append((rd + wr, us, tp))
        for name, v in acc.items():
            table[f"{name}@{cfg}"] = sum(a for a, _, _ in v) / len(v)
            meta[f"{name}@{cfg}"] = {"launches": len(v), "avg_duration_" + unit[ix["gpu__time_duration.sum"]]: sum(b for _, b, _ in v) / len(v),
                                     "tensor_pipe_pct": [c for _, _, c in v][:4], "source": os.path.basename(rep)}
    table["_meta"] = meta
    json.dump(table, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path, len(table) - 1, "entries")


if __name__ == "__main__":
    main()
