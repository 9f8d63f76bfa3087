#!/usr/bin/env python
"""Turn the ncu launch list of one bench step (the `--metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum,...` pass, see the command in the written file) into a markdown table + profiles/r2_traffic.json, which
bench.py reads for roofline.traffic.

    python tools/summarize_ncu_launches.py gpurun_out/launches.csv profiles/r2_ncu_launches_vN.md
"""
import csv
import json
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0] != "ID"]
    launches = OrderedDict()
    for r in rows:
        key = int(r[0])
        name = r[4].replace("void w2c::<unnamed>::", "").replace("w2c::<unnamed>::", "").split("(")[0]
        launches.setdefault(key, {"name": name, "grid": r[8], "block": r[7]})[r[-3]] = float(r[-1].replace(",", ""))
    total = sum(v["gpu__time_duration.sum"] for v in launches.values())
    conv = [v for v in launches.values() if v["name"].startswith("conv_") or v["name"].startswith("enc_head")]
    conv_bytes = sum(v.get("dram__bytes_read.sum", 0) + v.get("dram__bytes_write.sum", 0) for v in conv)
    conv_ns = sum(v["gpu__time_duration.sum"] for v in conv)
    with open(dst, "w") as f:
        f.write("# ncu launch list of one step (`python bench.py --profile-step`, 40 agent-frames, bf16)\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
                "sm__pipe_tensor_cycles_active...,lts__t_bytes.sum --clock-control none --profile-from-start off`; "
                "times are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| # | kernel | grid | us | share | DRAM rd MB | DRAM wr MB | L2 MB | tensor-pipe active % |\n")
        f.write("|---|---|---|---|---|---|---|---|---|\n")
        for k, v in launches.items():
            t = v["gpu__time_duration.sum"]
            f.write("| %d | %s | %s | %.1f | %.1f%% | %.1f | %.1f | %.1f | %.1f |\n" % (
                k, v["name"], v["grid"], t / 1e3, 100 * t / total, v.get("dram__bytes_read.sum", 0) / 1e6,
                v.get("dram__bytes_write.sum", 0) / 1e6, v.get("lts__t_bytes.sum", 0) / 1e6,
                v.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0)))
        f.write("\ntotal %.1f us over %d launches; tensor-core conv launches: %d, %.1f us (%.1f%% of the step), "
                "DRAM traffic %.1f MB per step = %.1f MB per launch\n" % (
                    total / 1e3, len(launches), len(conv), conv_ns / 1e3, 100 * conv_ns / total, conv_bytes / 1e6,
                    conv_bytes / 1e6 / max(1, len(conv))))
    sys.path.insert(0, ROOT)
    from multiagentperception_b200 import build as _b
    with open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w") as f:
        # csrc_fingerprint: bench.py reports this traffic only while the kernel sources are the ones it was measured on
        json.dump({"conv_launches": len(conv), "conv_dram_bytes_per_step": conv_bytes,
                   "conv_share_of_step_ncu": conv_ns / total, "source": os.path.basename(dst),
                   "csrc_fingerprint": _b._fingerprint(())[:16]}, f, indent=1)
    print("wrote", dst)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
