"""Build-container tool: per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`).
    python -m tests.tools.launch_summary gpurun_out/<launches>.csv [skip_first_n]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        us = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)
        rows.append((r[ik], us))
    rows = rows[skip:]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for k, us in rows:
        k = re.sub(r"\(.*", "", k)
        k = re.sub(r"^void |\(anonymous namespace\)::", "", k)
        tot[k][0] += 1
        tot[k][1] += us
    total = sum(v[1] for v in tot.values())
    print(f"{len(rows)} launches, {total / 1000:.2f} ms under ncu (serialised, cold cache: compare shares)")
    print("| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {us / 1000:.2f} | {us / n:.1f} | {100 * us / total:.1f}% |")


if __name__ == "__main__":
    main()
