"""Per-kernel stall picture from an ncu report's source page (read here, no GPU needed): totals per stall reason and the
top SASS lines with their dominant reasons.  usage: python tools/ncu_stalls.py report.ncu-rep [kernel-regex] [top]"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
regex = sys.argv[2] if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
extra = ["--kernel-name", f"regex:{regex}"] if regex else []
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + extra, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
i = 0
seen = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name" and i + 1 < len(rows) and "stall_long_sb" in rows[i + 1]:
        name = rows[i][1][:90]
        h = rows[i + 1]
        reasons = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
        j = i + 2
        tot = Counter()
        lines = []
        insts = 0
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            r = rows[j]
            j += 1
            if len(r) < len(h):
                continue
            try:
                smp = int(r[h.index("# Samples")])
                insts += int(r[h.index("Instructions Executed")])
            except ValueError:
                continue
            rs = {}
            for c in reasons:
                try:
                    v = int(r[h.index(c)])
                except ValueError:
                    v = 0
                if v:
                    rs[c[6:]] = v
                    tot[c[6:]] += v
            lines.append((smp, r[h.index("Source")].strip()[:64], rs, len(lines)))
        total = sum(l[0] for l in lines)
        print(f"===== launch {seen}: {name}\n  warp instructions {insts}, samples {total}")
        print("  reasons:", ", ".join(f"{k} {100 * v / max(1, sum(tot.values())):.1f}%" for k, v in tot.most_common(8)))
        for smp, src, rs, idx in sorted(lines, key=lambda l: -l[0])[:top]:
            main = ", ".join(f"{k}:{v}" for k, v in sorted(rs.items(), key=lambda kv: -kv[1])[:2])
            print(f"  {100 * smp / max(1, total):5.2f}%  #{idx:4d} {src:64s} {main}")
        seen += 1
        i = j
    else:
        i += 1
