"""Turn gpurun_out/{launches.csv, *.ncu-rep, bench_*.json} into the small text summaries kept under profiles/.
usage: python tools/summarize_profiles.py <tag> [--rep gpurun_out/prof_tc.ncu-rep] [--launches gpurun_out/launches.csv] [--bench f.json]"""
import argparse, collections, csv, io, json, os, subprocess, sys

KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max"]

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--rep")
    ap.add_argument("--launches")
    ap.add_argument("--bench")
    a = ap.parse_args()
    os.makedirs("profiles", exist_ok=True)
    if a.launches:
        rows = [r for r in csv.reader(open(a.launches)) if len(r) > 5]
        ci = {h: i for i, h in enumerate(rows[0])}
        agg = collections.defaultdict(lambda: [0, 0.0])
        for r in rows[1:]:
            try:
                v = float(r[ci["Metric Value"]].replace(",", ""))
            except ValueError:
                continue
            k = r[ci["Kernel Name"]].split("(")[0][:70]
            agg[k][0] += 1
            agg[k][1] += v
        tot = sum(v[1] for v in agg.values())
        with open(f"profiles/{a.tag}_launches.txt", "w") as f:
            f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n# source: {a.launches}\n")
            f.write(f"{'kernel':70s} {'launches':>8s} {'total_us':>12s} {'share':>7s}\n")
            for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
                f.write(f"{k:70s} {v[0]:8d} {v[1]/1000:12.1f} {v[1]/tot*100:6.1f}%\n")
        print(open(f"profiles/{a.tag}_launches.txt").read())
    if a.rep:
        raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        with open(f"profiles/{a.tag}_ncu_full.txt", "w") as f:
            f.write(f"# ncu --set full --clock-control none, source: {a.rep}\n")
            for r in rows[2:]:
                f.write("\n" + r[hdr.index("Kernel Name")][:100] + "\n")
                for i, h in enumerate(hdr):
                    if h in KEYS:
                        f.write(f"  {h:70s} {r[i]:>16s} {units[i]}\n")
        print(open(f"profiles/{a.tag}_ncu_full.txt").read()[:3000])
    if a.bench:
        j = json.loads(open(a.bench).read().strip().splitlines()[-1])
        json.dump(j, open(f"profiles/{a.tag}_bench.json", "w"), indent=1)

if __name__ == "__main__":
    main()
