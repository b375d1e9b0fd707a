"""Summarise an .ncu-rep (ncu --set full) into profiles/<name>.md: per kernel the duration, DRAM
bytes, L2/L1 hit rates, occupancy, issue utilisation and the top stall reasons.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name.md [note]"""
import csv
import io
import json
import os
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
        "sm__inst_executed.sum", "sm__cycles_elapsed.max"]
lines = ["# ncu summary: %s" % os.path.basename(rep), "", note, ""]
traffic = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0]
    lines.append("## %s" % name)
    lines.append("")
    lines.append("| metric | value | unit |")
    lines.append("|---|---|---|")
    vals = {}
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            lines.append("| %s | %s | %s |" % (w, r[i], units[i]))
            vals[w] = (r[i], units[i])
    def tobytes(v, u):
        f = float(v)
        return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
    if "dram__bytes_read.sum" in vals:
        traffic[name] = tobytes(*vals["dram__bytes_read.sum"]) + tobytes(*vals["dram__bytes_write.sum"])
        lines.append("| dram bytes read+write per launch | %.4g | byte |" % traffic[name])
    st = [(float(r[i]), h) for i, h in enumerate(hdr)
          if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")
          and r[i] not in ("", "n/a")]
    lines.append("")
    lines.append("top stall reasons (warps stalled per issue-active cycle): " + ", ".join(
        "%s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v)
        for v, h in sorted(st, reverse=True)[:6]))
    lines.append("")
open(out, "w").write("\n".join(lines) + "\n")
print(json.dumps(traffic))
