#!/usr/bin/env python3
"""Record the DRAM traffic of the vote kernel from an `ncu --set full` report into profiles/r02_traffic.json,
the file bench.py loads `roofline.traffic` from (instead of a literal in the source).

usage: tools/ncu_traffic.py REPORT.ncu-rep KEYFRAMES QUERIES N_GPUS [kernel-regex]
The record is keyed by kernel name + workload + sharding and carries the sha of search.cu at capture time,
so bench.py can say whether the kernel changed since (roofline.traffic_source.stale_vs_current_source)."""
import csv
import hashlib
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, nkf, nq, ngpu = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    pat = re.compile(sys.argv[5] if len(sys.argv) > 5 else r"k_vote_run|k_vote_join|k_vote\b")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]

    def val(r, name):
        i = hdr.index(name)
        v = float(r[i].replace(",", ""))
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6,
                 "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3}.get(units[i], 1.0)
        return v * scale

    sha = hashlib.sha256(open(os.path.join(ROOT, "sgtd_b200/csrc/search.cu"), "rb").read()).hexdigest()[:16]
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    doc = json.load(open(path)) if os.path.exists(path) else {"records": []}
    seen = set()
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("sgtd::", "")
        short = re.sub(r"<.*", "", name)
        if not pat.search(short) or short in seen:
            continue
        seen.add(short)
        rec = {"kernel": short, "keyframes": nkf, "queries": nq, "n_gpus": ngpu, "report": os.path.basename(rep),
               "source_sha": sha, "dram_read_bytes": val(r, "dram__bytes_read.sum"),
               "dram_write_bytes": val(r, "dram__bytes_write.sum"), "duration_ms": val(r, "gpu__time_duration.sum")}
        doc["records"] = [x for x in doc["records"] if (x["kernel"], x["keyframes"], x["queries"], x["n_gpus"]) !=
                          (short, nkf, nq, ngpu)] + [rec]
        print(rec)
    json.dump(doc, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
