"""Registers / spills / shared memory per kernel from the build log (lifcal_b200/build/*.o.log) -> short table for profiles/."""
import glob, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for log in sorted(glob.glob(os.path.join(ROOT, "lifcal_b200", "build", "*.o.log"))):
    name = None
    spill = ""
    for line in open(log):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = m.group(1)
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            spill = f"stack {m.group(1)} B, spill st/ld {m.group(2)}/{m.group(3)} B"
        m = re.search(r"Used (\d+) registers.*?(\d+) bytes smem", line)
        if m and name:
            try:
                dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            except OSError:
                dem = name
            dem = re.sub(r"\(.*", "", dem)
            rows.append((os.path.basename(log)[:-6], dem, int(m.group(1)), spill, m.group(2)))
            name, spill = None, ""
print(f"{'file':16s} {'kernel':70s} regs  static smem  spills")
for f, k, r, s, sm in rows:
    if "cub::" in k:
        continue
    print(f"{f:16s} {k[:70]:70s} {r:4d}  {sm:>8s} B   {s}")
