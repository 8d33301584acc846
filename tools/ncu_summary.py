"""Summarise `ncu --set full` captures of conv3x3_tc2 launches (one .ncu-rep per rung) into a markdown table and refresh
profiles/conv_traffic.json (DRAM bytes per launch, tied to the sha256 of the kernel source).  Runs without a GPU.
   python tools/ncu_summary.py --out profiles/r02_final_conv3x3_tc2_ncu_summary.md split=gpurun_out/a.ncu-rep fp16=gpurun_out/b.ncu-rep"""
import argparse
import csv
import hashlib
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    ("duration", "gpu__time_duration.sum"),
    ("SM active cycles", "sm__cycles_active.avg"),
    ("tensor pipe (mem view) active, % of elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("tensor pipe (mem view) active, % of active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("DRAM read", "dram__bytes_read.sum"),
    ("DRAM write", "dram__bytes_write.sum"),
    ("L2 -> SM bytes", "l1tex__m_xbar2l1tex_read_bytes.sum"),
    ("L2 sectors requested by SMs", "lts__t_sectors_srcunit_tex.sum"),
    ("registers / thread", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("dynamic shared memory / CTA", "launch__shared_mem_per_block_dynamic"),
    ("tcgen05.ld instructions", "smsp__sass_inst_executed_op_tmem_ldt.sum"),
    ("warp instructions executed", "smsp__inst_executed.sum"),
    ("issue slots busy, %", "sm__inst_issued.avg.pct_of_peak_sustained_active"),
]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        if len(r) == len(hdr):
            res.append({h: (v, u) for h, v, u in zip(hdr, r, units)})
    return res


ap = argparse.ArgumentParser()
ap.add_argument("--out", required=True)
ap.add_argument("--title", default="conv3x3_tc2 under `ncu --set full --clock-control none`")
ap.add_argument("--note", default="")
ap.add_argument("--traffic", action="store_true", help="refresh profiles/conv_traffic.json from the first launch of each capture")
ap.add_argument("captures", nargs="+", help="label=path.ncu-rep")
a = ap.parse_args()
cols = []
for c in a.captures:
    label, path = c.split("=", 1)
    for i, launch in enumerate(raw(path)):
        name = launch.get("Kernel Name", ("?", ""))[0]
        cols.append(("%s #%d" % (label, i), name, launch))
src = os.path.join(ROOT, "sayuri_b200", "csrc", "conv3x3_tc2.cuh")
sha = hashlib.sha256(open(src, "rb").read()).hexdigest()
lines = ["# " + a.title, "", a.note, "", "Kernel source sha256 `%s`." % sha, "",
         "| metric | " + " | ".join(c[0] for c in cols) + " |", "|---|" + "---|" * len(cols),
         "| kernel | " + " | ".join("`%s`" % c[1][:44] for c in cols) + " |"]
for nice, key in METRICS:
    cells = []
    for _, _, l in cols:
        v, u = l.get(key, ("n/a", ""))
        cells.append(("%s %s" % (v, u)).strip())
    lines.append("| %s (`%s`) | %s |" % (nice, key, " | ".join(cells)))
open(a.out, "w").write("\n".join(lines) + "\n")
print("wrote", a.out)
if a.traffic:
    tpath = os.path.join(ROOT, "profiles", "conv_traffic.json")
    ent = json.load(open(tpath)) if os.path.exists(tpath) else {}
    seen = set()
    for label, name, l in cols:
        rung = "fp32_split" if label.startswith("split") else "fp16"
        if rung in seen:
            continue
        seen.add(rung)

        def num(key):
            v, u = l[key]
            f = float(v.replace(",", ""))
            return f * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
        ent[rung] = {"dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
                     "kernel_source_sha256": sha,
                     "source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one tower-conv launch (config 2, "
                               "one layer per launch: --option conv_chain=0), " + os.path.relpath(a.out, ROOT)}
    json.dump(ent, open(tpath, "w"), indent=1)
    print("refreshed", tpath)
