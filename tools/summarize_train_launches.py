#!/usr/bin/env python
"""ncu launch list of ONE training step (tools/profile_train_step.py under
`ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
 --clock-control none --csv`) -> a markdown table aggregated per kernel.

    python tools/summarize_train_launches.py gpurun_out/train_launches_v3.csv profiles/r2_train_step_launches.md "title"
"""
import csv
import sys
from collections import OrderedDict, defaultdict


def load(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0] != "ID"]
    launches = OrderedDict()
    for r in rows:
        name = r[4].replace("void w2c::<unnamed>::", "").replace("w2c::<unnamed>::", "").split("(")[0]
        launches.setdefault(int(r[0]), {"name": name, "grid": r[8]})[r[-3]] = float(r[-1].replace(",", ""))
    return launches


def main(src, dst, title):
    launches = load(src)
    total = sum(v["gpu__time_duration.sum"] for v in launches.values())
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for v in launches.values():
        a = agg[v["name"][:70]]
        a[0] += 1
        a[1] += v["gpu__time_duration.sum"]
        a[2] += v.get("dram__bytes_read.sum", 0) + v.get("dram__bytes_write.sum", 0)
    with open(dst, "w") as f:
        f.write("# %s\n\n" % title)
        f.write("`ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                "--clock-control none` around one step of `tools/profile_train_step.py` (eager launches, no CUDA graph). "
                "Times are serialised and cold-cache: compare SHARES.\n\n")
        f.write("total %.2f ms over %d launches\n\n" % (total / 1e6, len(launches)))
        f.write("| kernel | launches | us | share | DRAM MB | achieved TB/s |\n|---|---|---|---|---|---|\n")
        for name, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            if a[1] / total < 0.002:
                continue
            f.write("| `%s` | %d | %.1f | %.1f%% | %.1f | %.2f |\n" % (name, a[0], a[1] / 1e3, 100 * a[1] / total, a[2] / 1e6,
                                                                 a[2] / a[1] / 1e3))
    print("wrote", dst)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "ncu launch list of one training step")
