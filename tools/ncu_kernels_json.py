"""ncu report -> profiles/ncu_kernels.json: per kernel and launch the figures bench.py's roofline quotes
(duration, SM issue-slot utilisation, active lanes per warp instruction, FP64 pipe, DRAM bytes, warp instructions).

  python tools/ncu_kernels_json.py <capture id> <command line of the capture> report.ncu-rep [more.ncu-rep ...] > profiles/ncu_kernels.json
"""
import csv
import json
import re
import subprocess
import sys

capture, command, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
M = {"gpu__time_duration.sum": ("duration_us", 1e-3),      # ns
     "sm__inst_issued.avg.pct_of_peak_sustained_active": ("issue_slot_pct", 1.0),
     "smsp__thread_inst_executed_per_inst_executed.ratio": ("active_lanes", 1.0),
     "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": ("fp64_pipe_pct", 1.0),
     "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": ("fp32_fma_pipe_pct", 1.0),
     "sm__warps_active.avg.pct_of_peak_sustained_active": ("achieved_occupancy_pct", 1.0),
     "smsp__inst_executed.sum": ("warp_instructions", 1.0),
     "launch__registers_per_thread": ("registers", 1.0),
     "sm__cycles_active.avg": ("sm_cycles_active", 1.0), "sm__cycles_elapsed.avg": ("sm_cycles_elapsed", 1.0)}
UNIT_SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}
BYTE_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
kernels = {}
for rep in reps:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in out.splitlines() if not l.startswith("==")))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = re.sub(r"^void ", "", r[hdr.index("Kernel Name")]).split("(")[0].split("<")[0]
        e = {"launch_id": int(r[0]), "report": rep.split("/")[-1]}
        for key, (label, scale) in M.items():
            if key in hdr:
                i = hdr.index(key)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                if label == "duration_us":
                    v *= UNIT_SCALE.get(units[i], 1e-3)
                e[label] = v
        dram = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            if key in hdr:
                i = hdr.index(key)
                try:
                    dram += float(r[i].replace(",", "")) * BYTE_SCALE.get(units[i], 1.0)
                except ValueError:
                    pass
        e["dram_bytes"] = dram
        kernels.setdefault(name, {"launches": []})["launches"].append(e)
json.dump({"capture": capture, "command": command, "kernels": kernels}, sys.stdout, indent=1)
