"""Per-source-line instruction / sample totals of one kernel from an ncu report captured with
--import-source on (needs -lineinfo).  usage: ncu_source_hotspots.py report.ncu-rep kernel_substr [launch_no]"""
import collections, csv, subprocess, sys
rep, want = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
import os
out = open(os.environ["NCU_SRC_CSV"]).read() if os.environ.get("NCU_SRC_CSV") else subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True, cwd="/tmp").stdout
csv.field_size_limit(1 << 30)
rows = list(csv.reader(out.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Function Name":
        cur = {"name": r[1], "rows": [], "file": fpath}
        blocks.append(cur)
    elif r and r[0] == "File Path":
        fpath = r[1]
    elif r and r[0] == "Line No":
        cur["hdr"] = r
    elif cur is not None and len(r) > 8:
        cur["rows"].append(r)
sel = [b for b in blocks if want in b["name"]]
# one block per (launch, file): a launch's blocks list every file once, so a file that repeats starts the next launch
launches = collections.OrderedDict()
k, seen = -1, set()
for b in sel:
    if k < 0 or b["file"] in seen:
        k += 1
        seen = set()
    seen.add(b["file"])
    launches.setdefault(k, []).append(b)
tot = collections.Counter(); thr = collections.Counter(); smp = collections.Counter(); text = {}
for b in launches[which]:
    h = b["hdr"]; iI = h.index("Instructions Executed"); iT = h.index("Thread Instructions Executed"); iS = h.index("# Samples")
    fn = b["file"].split("/")[-1]
    line = None
    for r in b["rows"]:
        if r[0] != "":
            line = (fn, int(r[0])); text[line] = r[1]
        if r[2] != "" and line is not None:   # a SASS row
            num = lambda x: int(x) if x.strip().lstrip("-").isdigit() else 0
            tot[line] += num(r[iI]); thr[line] += num(r[iT]); smp[line] += num(r[iS])
T = sum(tot.values()); S = sum(smp.values())
print(f"kernel {want} launch {which}: warp instr {T}, samples {S}")
byfile = collections.Counter()
for (fn, ln), v in tot.items(): byfile[fn] += v
print("by file:", {k: f"{v / T * 100:.1f}%" for k, v in byfile.items()})
order = sorted(tot.items(), key=lambda kv: kv[0]) if os.environ.get("NCU_ALL_LINES") else tot.most_common(60)
for line, v in order:
    if os.environ.get("NCU_ALL_LINES") and v / T < 0.0005:
        continue
    print(f"{line[0]:14s}:{line[1]:4d} instr {v / T * 100:5.1f}%  samples {smp[line] / max(S,1) * 100:5.1f}%  thr/instr {thr[line] / max(v, 1):5.1f} | {text[line].strip()[:90]}")
