"""Summarise an ncu report of k_brr_iteration into profiles/ (developer tool).

usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/NAME.json "what was captured"
Writes the launch metrics the README quotes, the stall-reason shares and the shared-memory wavefront
counts per source region, and refreshes profiles/traffic.json (read by bench.py into roofline.traffic)."""
import csv
import io
import json
import subprocess
import sys

rep, out, what = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ("gpu__time_duration", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput", "launch__", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "sm__throughput",
        "sm__warps_active", "smsp__inst_executed.sum", "smsp__issue_active", "sm__cycles_elapsed.avg", "lts__t_sectors_srcunit_tex_op_read.sum")
metrics = {h: [v, u] for h, u, v in zip(hdr, units, vals) if h.startswith(keep)}

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
sh = srows[1]
ix = {h: i for i, h in enumerate(sh)}
data = srows[2:]


def f(r, k):
    try:
        return float(r[ix[k]])
    except (ValueError, KeyError, IndexError):
        return 0.0


tot = sum(f(r, "# Samples") for r in data) or 1.0
stalls = {h: round(sum(f(r, h) for r in data) / tot, 4) for h in sh if h.startswith("stall_") and "Not Issued" not in h}
stalls = {k: v for k, v in sorted(stalls.items(), key=lambda t: -t[1]) if v >= 0.005}
gathers = [r for r in data if r[ix["Source"]].strip().startswith("LDS.64") and f(r, "L1 Wavefronts Shared") > 0]
gw = sum(f(r, "L1 Wavefronts Shared") for r in gathers)
gi = sum(f(r, "L1 Wavefronts Shared Ideal") for r in gathers)
gn = sum(f(r, "Instructions Executed") for r in gathers)
summary = {
    "what": what,
    "metrics": metrics,
    "warp_stall_shares": stalls,
    "sass_instructions": len(data),
    "sass_instructions_executed": sum(1 for r in data if f(r, "Instructions Executed") > 0),
    "lds64": {"instructions": gn, "wavefronts": gw, "ideal_wavefronts": gi, "wavefronts_per_instruction": gw / gn if gn else None},
    "shared_wavefronts_total": sum(f(r, "L1 Wavefronts Shared") for r in data),
    "shared_wavefronts_ideal": sum(f(r, "L1 Wavefronts Shared Ideal") for r in data),
}
json.dump(summary, open(out, "w"), indent=1)


def num(k):
    v, u = metrics[k]
    x = float(v.replace(",", ""))
    return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)


traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
json.dump({"dram_bytes_per_launch": traffic, "source": f"{out} (dram__bytes_read.sum + dram__bytes_write.sum)"}, open("profiles/traffic.json", "w"))
print(json.dumps({"traffic": traffic, "stalls": stalls, "lds64": summary["lds64"]}, indent=1))
