"""Summarise `ncu --set full` captures (.ncu-rep) into a markdown table + DRAM traffic JSON.

    python scripts/ncu_summarise.py gpurun_out/a.ncu-rep [b.ncu-rep ...] --md profiles/r02_ncu_summary.md \
        --json profiles/ncu_traffic.json

Per captured launch: duration, DRAM bytes read + written, achieved DRAM bandwidth against the measured copy bandwidth
of MEASURED_PEAKS.json, SM busy, and the tensor-pipe utilisation (FP64 DMMA kernels: `sm__inst_executed_pipe_fp64`;
tcgen05 kernels: the tensor pipe cycles).  ncu times are cold-cache and serialised: use the SHARES, not the absolutes."""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_rd",
    "dram__bytes_write.sum": "dram_wr",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active": "dmma_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
}
UNIT = {"nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr = i
            break
    if hdr is None:
        return []
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: j for j, n in enumerate(names)}
    res = []
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        d = {"kernel": r[col["Kernel Name"]], "id": r[col["ID"]]}
        for m, key in WANT.items():
            if m in col and r[col[m]] not in ("", "n/a"):
                try:
                    v = float(r[col[m]].replace(",", ""))
                except ValueError:
                    continue
                d[key] = v * UNIT.get(units[col[m]], 1.0)
        res.append(d)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reps", nargs="+")
    ap.add_argument("--md")
    ap.add_argument("--json")
    ap.add_argument("--title", default="ncu --set full captures")
    a = ap.parse_args()
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:  # noqa: BLE001
        hbm = 6531.3
    lines = ["# %s" % a.title, "",
             "HBM denominator: %.0f GB/s (MEASURED_PEAKS.json).  ncu serialises launches and flushes caches between "
             "replays, so durations are upper bounds; `GB/s` = (DRAM read + write) / duration of that launch." % hbm, "",
             "| capture | kernel | grid x block | regs | duration (us) | DRAM rd (MB) | DRAM wr (MB) | GB/s | of HBM | "
             "L2 hit % | SM busy % | tensor pipe % | DMMA sub-pipe % | FP64 pipe % | warps active % |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    traffic = {}
    for rep in a.reps:
        for d in load(rep):
            dur = d.get("duration", 0.0)
            rd, wr = d.get("dram_rd", 0.0), d.get("dram_wr", 0.0)
            gbs = (rd + wr) / dur / 1e9 if dur else 0.0
            f = lambda k, fmt="%.1f": (fmt % d[k]) if k in d else "-"
            short = d["kernel"].split("(")[0][-70:]
            lines.append("| %s | `%s` | %s x %s | %s | %.1f | %.1f | %.1f | %.0f | %.2f | %s | %s | %s | %s | %s | %s |" % (
                os.path.basename(rep).replace(".ncu-rep", ""), short, f("grid", "%d"), f("block", "%d"), f("regs", "%d"),
                dur * 1e6, rd / 1e6, wr / 1e6, gbs, gbs / hbm, f("l2_hit"), f("sm_pct"), f("tensor_pct"), f("dmma_pct"),
                f("fp64_pct"), f("occupancy")))
            key = short.split("::")[-1].split("<")[0]
            best = traffic.get(key)
            if best is None or dur > best["duration_us"] * 1e-6:
                traffic[key] = {"dram_bytes_per_launch": rd + wr, "duration_us": dur * 1e6, "gbs": gbs,
                                "capture": os.path.basename(rep), "kernel": short}
    text = "\n".join(lines) + "\n"
    if a.md:
        open(a.md, "w").write(text)
    else:
        sys.stdout.write(text)
    if a.json:
        json.dump(traffic, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
