"""Stall-reason totals, instruction mix and the hottest SASS lines of a kernel from an .ncu-rep captured with
--import-source on (read on the CPU box): python tools/ncu_stalls.py rep [n_lines]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print("# kernel:", rows[0][1][:110])
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: sum(int(r[ix[h]]) for r in data) for h in stalls}
T = sum(tot.values())
n_inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
print(f"# warp instructions executed: {n_inst}; stall samples: {T}")
print("stall reasons (share of samples):", ", ".join(f"{h[6:]} {100 * v / T:.1f}%" for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:9]))
mix = collections.Counter()
for r in data:
    op = [o for o in r[ix["Source"]].split() if not o.startswith("@")][0].split(".")[0]
    mix[op] += int(r[ix["Instructions Executed"]])
print("instruction mix:", ", ".join(f"{op} {100 * c / n_inst:.1f}%" for op, c in mix.most_common(14)))
print("hottest SASS lines (samples, executed, instruction, top stall):")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:n_top]:
    s = sorted(((h, int(r[ix[h]])) for h in stalls), key=lambda kv: -kv[1])[0]
    print(f"  {r[ix['# Samples']]:>8} {r[ix['Instructions Executed']]:>12}  {r[ix['Source']].strip()[:64]:<64} {s[0][6:]}")
