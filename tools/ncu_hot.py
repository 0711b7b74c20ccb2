"""Print the hottest SASS instructions (by stall samples) of one kernel from an ncu report.
usage: python tools/ncu_hot.py report.ncu-rep <kernel regex> [launch index] [top n]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
print(lines[start - 1][:150])
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
r = [x for x in csv.DictReader(io.StringIO("\n".join(lines[start:end]))) if x.get("# Samples") not in (None, "")]
tot = sum(int(x["# Samples"]) for x in r)
stall_cols = [c for c in r[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
print("total samples", tot, "instructions", len(r))
agg = {c: sum(int(x[c]) for x in r) for c in stall_cols}
print("stall mix:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
idx = sorted(range(len(r)), key=lambda i: -int(r[i]["# Samples"]))[:top]
for i in sorted(idx):
    x = r[i]
    st = {c[6:]: int(x[c]) for c in stall_cols if int(x[c]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{i:5d} {int(x['# Samples']):7d} {100*int(x['# Samples'])/max(tot,1):5.1f}%  {x['Source'][:70]:70s} {st}")
