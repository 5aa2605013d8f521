#!/usr/bin/env python
"""profiles/traffic.json from `ncu --set full` reports: dram__bytes_read.sum + dram__bytes_write.sum per launch of each headline kernel.
   python tools/ncu_traffic.py tag=path.ncu-rep ...      (tag: k_analysis | k_perbin<8,LMS> | k_synthesis)
bench.py copies the figure of its dominant kernel into `roofline.traffic` and names this file as the source."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    names, units, vals = rows[0], rows[1], rows[-1]
    return {n: (v, u) for n, u, v in zip(names, units, vals)}


def main():
    res = {"_source": "ncu --set full --clock-control none, one launch per kernel at configs[1] size (tools/gpu_r02_*.sh); bytes per launch"}
    for arg in sys.argv[1:]:
        tag, rep = arg.split("=", 1)
        m = metrics(rep)
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v, u = m[k]
            tot += float(v.replace(",", "")) * UNIT[u]
        res[tag] = int(tot)
        res[tag + " duration_us_under_ncu"] = float(m["gpu__time_duration.sum"][0].replace(",", "")) * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}.get(m["gpu__time_duration.sum"][1], 1.0)
        res[tag + " report"] = os.path.basename(rep)
    json.dump(res, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
