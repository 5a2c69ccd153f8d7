"""Runs a command while sampling NVML power / SM clock (development aid): python tools/power_run.py <cmd...>"""
import subprocess, sys, threading, time
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
rows, stop = [], threading.Event()
def run():
    while not stop.is_set():
        rows.append((pynvml.nvmlDeviceGetPowerUsage(h) / 1e3, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
        stop.wait(0.02)
t = threading.Thread(target=run, daemon=True); t.start()
out = subprocess.run(sys.argv[1:], capture_output=True, text=True).stdout.strip()
stop.set(); t.join()
half = rows[len(rows) // 2:]
med = lambda v: sorted(v)[len(v) // 2]
print(f"{out}   | power {med([r[0] for r in half]):.0f} W  sm {med([r[1] for r in half])} MHz")
