#!/usr/bin/env python
"""Summarise an ncu report (read here, on the CPU box) into a markdown table + JSON for profiles/.

    python scripts/ncu_summary.py gpurun_out/prof_update.ncu-rep profiles/r01_update_kernels
"""
import csv
import json
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1_pct"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__inst_executed.sum", "warp_insts"), ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts")]

TO_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TO_MS = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    recs = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")}
        for col, key in COLS:
            if col not in hdr:
                continue
            i = hdr.index(col)
            v = float(r[i].replace(",", "")) if r[i] else None
            u = units[i]
            if v is not None and u in TO_BYTES:
                v *= TO_BYTES[u]
            if v is not None and key == "time":
                v *= TO_MS.get(u, 1.0)
            d[key] = v
        d["dram_bytes"] = (d.get("dram_read") or 0) + (d.get("dram_write") or 0)
        d["dram_gbs"] = d["dram_bytes"] / (d["time"] * 1e-3) / 1e9
        recs.append(d)
    with open(out + ".json", "w") as fh:
        json.dump(recs, fh, indent=1)
    with open(out + ".md", "w") as fh:
        fh.write(f"ncu --set full --clock-control none, report `{rep}` (per launch; cold caches, serialised)\n\n")
        fh.write("| kernel | ms | DRAM read MB | DRAM write MB | DRAM GB/s | DRAM % | SM % | L1 % | L2 % | occ % | regs | grid x block | warp insts | smem wavefronts |\n")
        fh.write("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for d in recs:
            fh.write("| {kernel} | {time:.3f} | {r:.1f} | {w:.1f} | {g:.0f} | {dp:.1f} | {sp:.1f} | {l1:.1f} | {l2:.1f} | {oc:.1f} | {regs:.0f} | {grid:.0f} x {block:.0f} | {wi:.3g} | {sw:.3g} |\n".format(
                r=d["dram_read"] / 1e6, w=d["dram_write"] / 1e6, g=d["dram_gbs"], dp=d.get("dram_pct") or 0, sp=d.get("sm_pct") or 0,
                l1=d.get("l1_pct") or 0, l2=d.get("l2_pct") or 0, oc=d.get("occupancy_pct") or 0, wi=d.get("warp_insts") or 0,
                sw=d.get("smem_wavefronts") or 0, **{k: d[k] for k in ("kernel", "time", "regs", "grid", "block")}))
    print(open(out + ".md").read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
