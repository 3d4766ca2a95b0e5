"""Turn an `ncu --set full` report into the per-kernel facts bench.py prints next to its live timings.

    python scripts/ncu_traffic.py <report.ncu-rep> <channels> <samples> [--mix-ceiling IPC --mix-source TEXT]

For every kernel family bench.py knows (bench.KERNEL_SOURCES) the first matching launch in the report
gives: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) and executed warp instructions
(smsp__inst_executed.sum), both divided by channels x samples of that launch, plus the sha256 of the
kernel's source file (comments and white space removed) AS IT IS NOW -- run this right after the capture, on the tree that was captured.
bench.py refuses the record (prints traffic: null, "stale") once the source file changes.
Writes profiles/r02_traffic.json (merging with what is there).
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    return rows[2:], idx, units


def num(s):
    return float(s.replace(",", ""))


def scale(unit):
    unit = unit.lower()
    return {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "inst": 1.0}.get(unit, 1.0)


def main():
    path, C, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    mix, mix_src = None, None
    if "--mix-ceiling" in sys.argv:
        mix = float(sys.argv[sys.argv.index("--mix-ceiling") + 1])
        mix_src = sys.argv[sys.argv.index("--mix-source") + 1] if "--mix-source" in sys.argv else None
    rows, idx, units = rows_of(path)
    out_path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    rec = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for kernel, src in bench.KERNEL_SOURCES.items():
        hit = next((r for r in rows if kernel in r[idx["Kernel Name"]]), None)
        if hit is None:
            continue
        rd = num(hit[idx["dram__bytes_read.sum"]]) * scale(units[idx["dram__bytes_read.sum"]])
        wr = num(hit[idx["dram__bytes_write.sum"]]) * scale(units[idx["dram__bytes_write.sum"]])
        inst = num(hit[idx["smsp__inst_executed.sum"]])
        dur = num(hit[idx["gpu__time_duration.sum"]])
        e = {"dram_bytes_per_channel_sample": (rd + wr) / (C * T),
             "warp_inst_per_channel_sample": inst / (C * T),
             "source_sha": bench.source_sha(src),
             "source": f"ncu --set full --clock-control none, {hit[idx['Kernel Name']][:60]} at {C} ch x {T} samples "
                       f"({os.path.basename(path)}): dram read {rd / 1e9:.3f} GB + write {wr / 1e9:.3f} GB, "
                       f"{inst / 1e9:.3f} G warp instructions, {dur} {units[idx['gpu__time_duration.sum']]} under ncu"}
        if mix is not None and kernel == "hilbert_env8_kernel":
            e["mix_ceiling_ipc"] = mix
            e["mix_ceiling_source"] = mix_src
        rec[kernel] = e
        print(kernel, json.dumps(e, indent=1))
    with open(out_path, "w") as f:
        json.dump(rec, f, indent=1)


if __name__ == "__main__":
    main()
