"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table.
    python tools/summarize_launches.py gpurun_out/launches_r1e.csv profiles/launches_r1e.md "<command>" """
import collections
import csv
import sys


def main(src, dst, command):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    d = collections.OrderedDict()
    for x in csv.DictReader(lines):
        try:
            v = float(x["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = x.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        d.setdefault(x["Kernel Name"], []).append(v)
    tot = sum(sum(v) for v in d.values())
    out = [f"# Launch list ({src.split('/')[-1]}; ncu --metrics gpu__time_duration.sum --clock-control none)", "",
           f"Command: `{command}` (bench under ncu: cold-cache, serialised launches -> compare SHARES, not absolutes).", "",
           "| kernel | launches | total us | avg us | share of all GPU time |", "|---|---|---|---|---|"]
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        if sum(v) / tot < 0.002:
            continue
        out.append(f"| `{k[:90]}` | {len(v)} | {sum(v):.1f} | {sum(v) / len(v):.2f} | {100 * sum(v) / tot:.1f}% |")
    open(dst, "w").write("\n".join(out) + "\n")
    print("wrote", dst)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
