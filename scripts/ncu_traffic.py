"""Summarise `ncu --set full` captures (gpurun_out/*.ncu-rep) into tracked files under profiles/:

  profiles/ncu_traffic.json      {key: {"dram_bytes": read + write per launch, "source": file, ...}}  (read by bench.py's
                                 roofline.traffic -- the number is never a literal in bench.py)
  profiles/<name>_summary.txt    duration, cycles, tensor / XU / issue utilisation, DRAM bytes and the warp-stall mix

Usage: python scripts/ncu_traffic.py key=path.ncu-rep[:round_tag] ...   (runs here, no GPU: ncu only reads the report)
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg", "sm cycles"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe % of peak (elapsed)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % of peak (active)"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "XU (MUFU) pipe %"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue slots busy %"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA-heavy pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__bytes_read.sum.per_second", "dram read rate"),
    ("dram__bytes_write.sum.per_second", "dram write rate"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers)"),
]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def ncu_csv(rep, page):
    r = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr[-2000:])
    return list(csv.reader(io.StringIO(r.stdout)))


def summarise(key, rep, tag):
    rows = ncu_csv(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    name = vals[col["Kernel Name"]] if "Kernel Name" in col else key
    lines = [f"# {os.path.basename(rep)}  ({tag})", f"kernel: {name}", ""]
    got = {}
    for m, label in METRICS:
        if m in col and vals[col[m]] != "":
            got[m] = (float(vals[col[m]].replace(",", "")), units[col[m]])
            lines.append(f"{label:38s} {vals[col[m]]:>16s} {units[col[m]]}")
    rd, ru = got.get("dram__bytes_read.sum", (0.0, "byte"))
    wr, wu = got.get("dram__bytes_write.sum", (0.0, "byte"))
    dram = rd * UNIT.get(ru, 1.0) + wr * UNIT.get(wu, 1.0)
    dur, du = got.get("gpu__time_duration.sum", (0.0, "us"))
    dur_s = dur * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(du, 1e-6)
    if dur_s > 0:
        lines.append(f"{'dram bytes per launch':38s} {dram / 1e6:16.1f} MB  -> {dram / dur_s / 1e9:.0f} GB/s achieved")
    # warp-stall mix from the source page (sampling): all samples, by reason
    try:
        src = ncu_csv(rep, "source")
        h2 = src[1]
        ix = {h: i for i, h in enumerate(h2)}
        stalls = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
        tot = {h: 0 for h in stalls}
        n = 0
        for r in src[2:]:
            n += int(r[ix["# Samples"]])
            for h in stalls:
                tot[h] += int(r[ix[h]])
        lines += ["", f"warp-stall sampling, all warps ({n} samples):"]
        for h, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            if v:
                lines.append(f"  {h:26s} {100.0 * v / max(n, 1):5.1f} %")
    except Exception as e:  # a capture without the source counters still gives the raw page
        lines.append(f"(no source page: {e})")
    out = os.path.join(ROOT, "profiles", f"{tag}_ncu_{key}_summary.txt")
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    return {"dram_bytes": dram, "source": f"profiles/{os.path.basename(out)}", "kernel": name,
            "duration_us": dur_s * 1e6}


def main():
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        table = json.load(open(path))
    except (OSError, ValueError):
        table = {}
    for arg in sys.argv[1:]:
        key, rest = arg.split("=", 1)
        rep, _, tag = rest.partition(":")
        table[key] = summarise(key, rep, tag or "r2")
        print(key, table[key])
    with open(path, "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)
        f.write("\n")


if __name__ == "__main__":
    main()
