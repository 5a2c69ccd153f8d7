"""Prints the last N launches of an `ncu --csv` metrics log as a table (development aid)."""
import collections
import csv
import io
import sys

path, last = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 20
lines = open(path).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
by = collections.OrderedDict()
for r in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
    d = by.setdefault(int(r["ID"]), {"name": r["Kernel Name"].split("(")[0][-34:], "grid": r["Grid Size"], "block": r["Block Size"]})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
tot = 0.0
for k in list(by)[-last:]:
    d = by[k]
    t = d.get("gpu__time_duration.sum", 0)
    tot += t
    rd, wr = d.get("dram__bytes_read.sum", 0), d.get("dram__bytes_write.sum", 0)
    fma, inst = d.get("sm__inst_executed_pipe_fma.sum", 0), d.get("sm__inst_executed.sum", 0)
    tp = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    extra = f" fma-share {fma / inst:.2f}  fma-rate {fma * 32 / t / 1e3:.1f} TFMA/s(issued)" if inst else ""
    extra += f" tensor {tp:.1f}%" if tp is not None else ""
    print(f"{k:4d} {d['name']:34s} {d['grid']:>16s} {t / 1e3:7.1f} us  rd {rd / 1e6:7.2f} wr {wr / 1e6:7.2f} MB{extra}")
print(f"sum {tot / 1e3:.1f} us")
