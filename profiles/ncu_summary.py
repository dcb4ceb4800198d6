"""Condense `ncu -i X.ncu-rep --page raw --csv` into the per-launch figures the bench line and the notes quote.

    ncu -i profiles/r2_head4_full.ncu-rep --page raw --csv > /tmp/raw.csv
    python profiles/ncu_summary.py /tmp/raw.csv "<command line that was captured>" > profiles/head_kernel_ncu.json
"""
import csv
import json
import sys

_STALL = "smsp__average_warps_issue_stalled_"


def _num(cell):
    try:
        return float(cell.replace(",", ""))
    except ValueError:
        return None


def summarise(path, source=""):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}

    def get(row, name):
        v = _num(row[col[name]])
        return None if v is None else v * scale.get(units[col[name]], 1.0)

    launches = []
    for r in data:
        stalls = {h[len(_STALL):-len("_per_issue_active.ratio")]: round(_num(r[i]), 3) for h, i in col.items()
                  if h.startswith(_STALL) and h.endswith("_per_issue_active.ratio") and (_num(r[i]) or 0) >= 0.04}
        launches.append({
            "duration_us": get(r, "gpu__time_duration.sum"),
            "dram_read_bytes": get(r, "dram__bytes_read.sum"),
            "dram_write_bytes": get(r, "dram__bytes_write.sum"),
            "grid": int(get(r, "launch__grid_size")), "block": int(get(r, "launch__block_size")),
            "regs": int(get(r, "launch__registers_per_thread")),
            "issue_active_pct": get(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") if "smsp__issue_active.avg.pct_of_peak_sustained_active" in col else None,
            "tensor_pipe_active_pct": next((get(r, h) for h in hdr if h.startswith("sm__pipe_tensor") and h.endswith("cycles_active.avg.pct_of_peak_sustained_active")), None),
            "warp_inst": get(r, "smsp__inst_executed.sum"),
            "l1_sector_hit_pct": get(r, "l1tex__t_sector_hit_rate.pct"),
            "lts_sectors": get(r, "lts__t_sectors.sum"),
            "stall_per_issue": stalls,
        })
    kernel = data[0][col["Kernel Name"]] if "Kernel Name" in col else ""
    traffic = sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in launches) / len(launches)
    return {"source": source, "kernel": kernel, "launches": launches, "dram_bytes_per_launch": traffic}


if __name__ == "__main__":
    print(json.dumps(summarise(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""), indent=1))
