"""Per-launch table of the LAST Generator + Detector pass in an ncu launch list of tools/m14_run.py:
python tools/m14_launch_table.py gpurun_out/m14_launches.csv [passes]"""
import csv
import sys

lines = [l for l in open(sys.argv[1], errors="replace") if not l.startswith("==")]
L = []
for x in csv.DictReader(lines):
    v = float(x["Metric Value"].replace(",", ""))
    u = x["Metric Unit"]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    L.append((int(x["ID"]), x["Kernel Name"].replace("void ", "").replace("wm::<unnamed>::", "").split("(")[0][:48], ms, x.get("Grid Size")))
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = len(L) // passes
last = L[-n:]
print("launches per pass", n, "sum ms", round(sum(x[2] for x in last), 3))
for x in last:
    if x[2] >= 0.02:
        print(x[0], x[1].ljust(48), "%.3f" % x[2], x[3])
