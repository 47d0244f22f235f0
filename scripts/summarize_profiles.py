"""Turn the raw outputs of scripts/collect_profiles.sh (gpurun_out/) into the tracked files under profiles/:
bench JSON lines, the launch-list table, per-GEMM DRAM traffic (gemm_traffic.json), kbench text and the --set full summaries.

    python scripts/summarize_profiles.py [round-tag, default r01]

Hand-written analysis (the .md sections around the tables) is not touched: tables are written to profiles/<tag>_auto_*.md and the
JSON / CSV artefacts are refreshed in place."""
import csv
import gzip
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
LAYERS = 27  # so400m


def copy(src, dst):
    s = os.path.join(O, src)
    if os.path.exists(s):
        shutil.copyfile(s, os.path.join(P, dst))
        print("copied", src, "->", dst)


def gz(src, dst):
    s = os.path.join(O, src)
    if os.path.exists(s):
        with open(s, "rb") as f, gzip.open(os.path.join(P, dst), "wb") as g:
            shutil.copyfileobj(f, g)
        print("gzipped", src, "->", dst)


def traffic():
    path = os.path.join(O, f"{TAG}_gemm_dram.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    mn, mv, mu, idc = (hdr.index(k) for k in ("Metric Name", "Metric Value", "Metric Unit", "ID"))
    per = {}
    for r in rows[hi + 1:]:
        if len(r) != len(hdr):
            continue
        v = float(r[mv].replace(",", ""))
        if "byte" in r[mu].lower():
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[mu]]
        per.setdefault(r[idc], {})[r[mn]] = v
    ids = sorted(per, key=int)
    rd = sum(per[i]["dram__bytes_read.sum"] for i in ids)
    wr = sum(per[i]["dram__bytes_write.sum"] for i in ids)

    def tot(i):
        return per[ids[i]]["dram__bytes_read.sum"] + per[ids[i]]["dram__bytes_write.sum"]

    names = ["qkv", "out+res", "fc1", "fc2+res"]  # launch 0 is the patch embedding, then 4 GEMMs per encoder layer
    meas = {n: sum(tot(1 + 4 * l + j) for l in range(LAYERS)) / LAYERS for j, n in enumerate(names)}
    ms = {n: sum(per[ids[1 + 4 * l + j]]["gpu__time_duration.sum"] for l in range(LAYERS)) / LAYERS / 1e6 for j, n in enumerate(names)}
    out = os.path.join(P, f"{TAG}_gemm_traffic.json")
    old = json.load(open(out)) if os.path.exists(out) else {}
    old.update({"dram_bytes_per_launch": (rd + wr) / len(ids), "launches": len(ids), "dram_read_bytes_per_step": rd,
                "dram_write_bytes_per_step": wr, "measured_bytes_per_layer": meas, "ms_under_ncu_per_layer": ms})
    json.dump(old, open(out, "w"), indent=1)
    print("gemm_traffic.json:", round((rd + wr) / len(ids) / 1e9, 3), "GB per launch", {k: round(v / 1e9, 2) for k, v in meas.items()})


def table(kind, src, dst):
    s = os.path.join(O, src)
    if not os.path.exists(s):
        return
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), kind, s], capture_output=True, text=True).stdout
    open(os.path.join(P, dst), "w").write(txt)
    print("wrote", dst)


if __name__ == "__main__":
    # scripts/collect_profiles.sh <tag> writes gpurun_out/<tag>_*: bench lines, kbench text and micro-benchmarks are copied as
    # they are, launch lists are compressed, the --set full reports are reduced to their key metrics
    for f in sorted(os.listdir(O)):
        if not f.startswith(TAG + "_"):
            continue
        if f.endswith((".json", ".txt")) and os.path.getsize(os.path.join(O, f)) > 0:
            copy(f, f)
    gz(f"{TAG}_launches_bench.csv", f"{TAG}_launches_bench.csv.gz")
    gz(f"{TAG}_gemm_dram.csv", f"{TAG}_gemm_dram.csv.gz")
    traffic()
    table("launches", f"{TAG}_launches_bench.csv", f"{TAG}_auto_launch_table.md")
    table("full", f"{TAG}_prof_bench_gemm.ncu-rep", f"{TAG}_auto_gemm_full.md")
    table("full", f"{TAG}_prof_bench_attn.ncu-rep", f"{TAG}_auto_attention_full.md")
