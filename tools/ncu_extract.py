"""CPU-side reader of the ncu reports brought back in gpurun_out/: per captured launch the duration, DRAM bytes,
tensor-pipe / DRAM utilisation, occupancy and registers -> profiles/r02_ncu_summary.md (table) and
profiles/r02_ncu_traffic.json ({kernel: dram bytes per launch}, read by bench.py for `roofline.traffic`).
    python tools/ncu_extract.py gpurun_out/r02_prof_probe.ncu-rep gpurun_out/r02_prof_vae.ncu-rep ...
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {
    "gpu__time_duration.sum": "dur",
    "dram__bytes_read.sum": "rd",
    "dram__bytes_write.sum": "wr",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct2",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__inst_executed.sum": "inst",
}
UNIT = {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def read(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: i for i, n in enumerate(names)}
    res = []
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        d = {"kernel": r[col["Kernel Name"]].split("(")[0].replace("void ", "").strip()}
        for metric, key in WANT.items():
            if metric in col and r[col[metric]] not in ("", "n/a"):
                v = float(r[col[metric]].replace(",", ""))
                d[key] = v * UNIT.get(units[col[metric]], 1.0)
        res.append(d)
    return res


def main():
    table, traffic = [], {}
    for rep in sys.argv[1:]:
        for d in read(rep):
            d["report"] = os.path.basename(rep)
            table.append(d)
    by = {}
    for d in table:
        by.setdefault(d["kernel"], []).append(d)
    lines = ["| kernel | launches | grid x block | regs | duration us (min / median) | DRAM read / write MB (median) | DRAM GB/s (median) | "
             "dram % | tensor pipe % (max) | warps active % |", "|---|---|---|---|---|---|---|---|---|---|"]
    med = lambda xs: sorted(xs)[len(xs) // 2]
    for k, ds in by.items():
        dur = [d.get("dur", 0.0) for d in ds]
        rd, wr = med([d.get("rd", 0.0) for d in ds]), med([d.get("wr", 0.0) for d in ds])
        gbs = med([(d.get("rd", 0.0) + d.get("wr", 0.0)) / max(d.get("dur", 1.0), 1e-9) / 1e3 for d in ds])
        tens = max(max(d.get("tensor_pct", 0.0), d.get("tensor_pct2", 0.0)) for d in ds)
        lines.append(f"| `{k}` | {len(ds)} | {int(ds[0].get('grid', 0))} x {int(ds[0].get('block', 0))} | {int(ds[0].get('regs', 0))} | "
                     f"{min(dur):.1f} / {med(dur):.1f} | {rd / 1e6:.2f} / {wr / 1e6:.2f} | {gbs:.0f} | "
                     f"{med([d.get('dram_pct', 0.0) for d in ds]):.0f} | {tens:.1f} | {med([d.get('warps_pct', 0.0) for d in ds]):.0f} |")
        traffic[k.split("<")[0]] = {"dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr, "duration_us_median": med(dur),
                                    "tensor_pipe_pct_max": tens, "launches_captured": len(ds),
                                    "dram_pct_median": med([d.get("dram_pct", 0.0) for d in ds])}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json"), "w") as fh:
        json.dump(traffic, fh, indent=1)
    # every captured launch, in capture order
    lines += ["", "| report | kernel | grid | duration us | DRAM read / write MB | tensor pipe % | dram % | inst executed |",
              "|---|---|---|---|---|---|---|---|"]
    for d in table:
        lines.append(f"| {d['report'].replace('.ncu-rep', '')} | `{d['kernel']}` | {int(d.get('grid', 0))} | {d.get('dur', 0.0):.1f} | "
                     f"{d.get('rd', 0.0) / 1e6:.2f} / {d.get('wr', 0.0) / 1e6:.2f} | "
                     f"{max(d.get('tensor_pct', 0.0), d.get('tensor_pct2', 0.0)):.1f} | {d.get('dram_pct', 0.0):.1f} | {int(d.get('inst', 0))} |")
    print("\n".join(lines))
    with open(os.path.join(ROOT, "profiles", "r02_ncu_table.md"), "w") as fh:
        fh.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
