#!/usr/bin/env python
"""Read an .ncu-rep here (no GPU needed): per-kernel top stall sites from the source page.

    python tools/ncu_hotspots.py gpurun_out/lenet_step.ncu-rep fc2_loss [top_n]
"""
import csv
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    print(lines[start - 1][:120])
    end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"') or not lines[i].strip()), len(lines))
    rows = [r for r in csv.DictReader(lines[start:end]) if r.get("# Samples") not in (None, "")]
    tot = sum(int(r["# Samples"]) for r in rows) or 1
    stall_cols = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
    agg = {c: sum(int(r[c] or 0) for r in rows) for c in stall_cols}
    print("samples=%d  instrs=%d" % (tot, len(rows)))
    print("stall mix: " + ", ".join("%s=%.0f%%" % (k[6:], 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    ranked = sorted(enumerate(rows), key=lambda ir: -int(ir[1]["# Samples"]))[:top]
    for i, r in sorted(ranked):
        s = int(r["# Samples"])
        why = max(stall_cols, key=lambda c: int(r[c] or 0))
        print("%5d %5.1f%%  %-14s exec=%-6s %s" % (i, 100.0 * s / tot, why[6:], r["Instructions Executed"], r["Source"].strip()[:110]))


if __name__ == "__main__":
    main()
