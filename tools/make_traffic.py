#!/usr/bin/env python
"""ncu per-launch metrics of ONE tagging call (tools/gpu_ncu_all.sh -> all_kernels.csv) -> per-kernel-class table
(profiles/<tag>_all_kernels.txt) and profiles/traffic.json, which bench.py reads for `roofline.traffic` (DRAM bytes of the
GEMM launches per clip) and for the measured DRAM bytes of the HBM-bound kernels.
usage: tools/make_traffic.py <all_kernels.csv> <tag> <model> <n_mels> <low 0|1> <clips>"""
import collections, csv, json, os, sys

src, tag, model, n_mels, low, clips = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
rows = [r for r in csv.reader(open(src)) if len(r) > 8]
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()            # launch id -> {metric: value}
names = {}
for r in rows[1:]:
    try:
        v = float(r[ci["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    lid = int(r[ci["ID"]])
    names[lid] = r[ci["Kernel Name"]]
    unit = r[ci["Metric Unit"]]
    m = r[ci["Metric Name"]]
    if m.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    if m == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(unit, 1)
    per.setdefault(lid, {})[m] = v


# keep exactly ONE tagging call: from the first fill_kernel (the mel front end starts every call with it) up to the next one
ids = sorted(per)
starts = [i for i in ids if "fill_kernel" in names[i]]
if len(starts) >= 2:
    per = collections.OrderedDict((i, per[i]) for i in ids if starts[0] <= i < starts[1])
elif len(starts) == 1:
    per = collections.OrderedDict((i, per[i]) for i in ids if i >= starts[0])


def klass(n):
    if "gemm_tc" in n: return "gemm"
    if "attn_tc" in n: return "attention"
    if "mel_" in n or "fill_kernel" in n: return "mel"
    if "pool20" in n: return "pool"
    if "layernorm" in n: return "layernorm"
    if "attn_small" in n: return "head_attention"
    return "other"


agg = collections.defaultdict(lambda: collections.defaultdict(float))
for lid, m in per.items():
    k = klass(names[lid])
    a = agg[k]
    a["launches"] += 1
    a["us"] += m.get("gpu__time_duration.sum", 0.0)
    a["dram"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    a["tensor_w"] += m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * m.get("gpu__time_duration.sum", 0.0)
tot = sum(a["us"] for a in agg.values())
os.makedirs("profiles", exist_ok=True)
with open(f"profiles/{tag}_all_kernels.txt", "w") as f:
    f.write(f"# ncu per-launch metrics of one tagging call ({model}, {clips} clips): cold-cache, serialised -> compare SHARES\n# source: {src}\n")
    f.write(f"{'class':16s} {'launches':>8s} {'total_us':>10s} {'share':>7s} {'dram_MB':>10s} {'tensor%':>8s}\n")
    for k, a in sorted(agg.items(), key=lambda x: -x[1]["us"]):
        f.write(f"{k:16s} {int(a['launches']):8d} {a['us']:10.1f} {100 * a['us'] / tot:6.1f}% {a['dram'] / 1e6:10.1f} {a['tensor_w'] / max(a['us'], 1e-9):8.1f}\n")
    f.write("\n# per kernel name\n")
    byname = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for lid, m in per.items():
        b = byname[names[lid].split("(")[0][:60]]
        b[0] += 1; b[1] += m.get("gpu__time_duration.sum", 0.0); b[2] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    for k, b in sorted(byname.items(), key=lambda x: -x[1][1]):
        f.write(f"{k:60s} {b[0]:5d} {b[1]:10.1f} us {b[2] / 1e6:10.1f} MB\n")
print(open(f"profiles/{tag}_all_kernels.txt").read())
tj = "profiles/traffic.json"
t = json.load(open(tj)) if os.path.exists(tj) else {}
t[f"{model}|{n_mels}|{low}"] = dict(source=f"profiles/{tag}_all_kernels.txt (ncu dram__bytes_read.sum + dram__bytes_write.sum, {clips} clips per call)",
                                    gemm_dram_bytes_per_clip=agg["gemm"]["dram"] / clips,
                                    dram_bytes_per_clip={k: agg[k]["dram"] / clips for k in ("mel", "pool", "layernorm", "attention") if k in agg})
json.dump(t, open(tj, "w"), indent=1)
print("traffic.json updated:", json.dumps(t[f"{model}|{n_mels}|{low}"])[:300])
