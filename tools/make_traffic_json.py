#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` capture of the bench command (run here, no GPU needed):

    python tools/make_traffic_json.py gpurun_out/<capture>.ncu-rep ["source note"]

Records dram__bytes_read.sum / dram__bytes_write.sum per launch for every kernel in the capture (the first launch
of each kernel name) together with the SHA-256 of the kernel sources the library was built from -- bench.py quotes
the numbers only while that hash still matches (a stale capture is reported as stale, not quoted)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep = sys.argv[1]
    note = sys.argv[2] if len(sys.argv) > 2 else "ncu --set full --clock-control none on `python bench.py`"
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = {"source": "%s (%s)" % (os.path.basename(rep), note), "kernel_sources_sha256": bench.kernel_sources_sha256()}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("<")[0].split("(")[0].split("::")[-1].replace("void ", "").strip()
        if name in res:
            continue
        entry = {}
        for key, col in (("dram_bytes_read", "dram__bytes_read.sum"), ("dram_bytes_write", "dram__bytes_write.sum"),
                         ("duration_us_under_ncu", "gpu__time_duration.sum")):
            v = float(r[idx[col]].replace(",", ""))
            u = units[idx[col]]
            if key.startswith("dram"):
                v *= UNITS[u]
            elif u == "ms":
                v *= 1e3
            elif u == "ns":
                v /= 1e3
            entry[key] = v
        entry["kernel"] = r[idx["Kernel Name"]][:120]
        res[name] = entry
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:   # the range-replay numbers of whole encodes (tools/range_traffic.py) are kept across per-kernel captures
        old = json.load(open(path))
        if "whole_encode_range_replay" in old:
            res["whole_encode_range_replay"] = old["whole_encode_range_replay"]
    except Exception:
        pass
    json.dump(res, open(path, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
