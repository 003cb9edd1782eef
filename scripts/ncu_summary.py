"""Per-kernel summary of an `ncu --page raw --csv` export: duration, DRAM bytes, tensor-pipe activity,
IPC, achieved occupancy, registers.    python scripts/ncu_summary.py RAW.csv [traffic.json]"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def num(r, name, scale=1.0):
    try:
        return float(r[col[name]].replace(",", "")) * scale
    except (KeyError, ValueError):
        return float("nan")


def to_bytes(r, name):
    u = units[col[name]].lower()
    s = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    return num(r, name, s)


def to_ms(r, name):
    u = units[col[name]].lower()
    s = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1)
    return num(r, name, s)


print(f"{'kernel':44s} {'grid':>7s} {'blk':>4s} {'regs':>4s} {'ms':>7s} {'dram rd MB':>10s} {'dram wr MB':>10s} "
      f"{'tensor %':>8s} {'IPC':>5s} {'warps %':>7s} {'L2 hit %':>8s}")
traffic = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    short = name.split("(")[0].replace("void ", "")[:44]
    rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
    print(f"{short:44s} {r[col['launch__grid_size']]:>7s} {r[col['launch__block_size']]:>4s} "
          f"{r[col['launch__registers_per_thread']]:>4s} {to_ms(r, 'gpu__time_duration.sum'):7.3f} "
          f"{rd / 1e6:10.1f} {wr / 1e6:10.1f} "
          f"{num(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):8.1f} "
          f"{num(r, 'sm__inst_executed.avg.per_cycle_elapsed'):5.2f} "
          f"{num(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):7.1f} "
          f"{num(r, 'lts__t_sector_hit_rate.pct'):8.1f}")
    key = short.split("<")[0]
    t = traffic.setdefault(key, {"read": 0.0, "write": 0.0, "launches": 0, "ms": 0.0})
    t["read"] += rd
    t["write"] += wr
    t["launches"] += 1
    t["ms"] += to_ms(r, "gpu__time_duration.sum")
if len(sys.argv) > 2:
    json.dump(traffic, open(sys.argv[2], "w"), indent=1)
