"""One line per launch from an `ncu --metrics ... --csv` log (tools/gpu_ncu_all.sh, tools/gpu_ncu_tail.sh).
usage: python tools/summarize_launch_counters.py <csv> <out.txt> [title]"""
import collections, csv, sys

def main():
    src, dst = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = [r for r in csv.reader(open(src)) if len(r) > 8]
    ci = {h: i for i, h in enumerate(rows[0])}
    L = collections.OrderedDict()
    for r in rows[1:]:
        d = L.setdefault(int(r[ci['ID']]), {'name': r[ci['Kernel Name']].split('(')[0].replace('void ', '').replace('wat::', '')})
        try:
            d[r[ci['Metric Name']]] = float(r[ci['Metric Value']].replace(',', ''))
        except ValueError:
            pass
        d['unit_' + r[ci['Metric Name']]] = r[ci['Metric Unit']]
    def by(d, k):
        return d.get(k, 0) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(d.get('unit_' + k, 'byte'), 1)
    out = [f"# {title}", "# ncu --metrics (per-launch counters), --clock-control none; cold-cache, serialised launches", ""]
    for i, d in L.items():
        if 'f32_to_bf16' in d['name']:
            continue
        t = d['gpu__time_duration.sum']
        t_us = t / 1000 if d['unit_gpu__time_duration.sum'] in ('ns', 'nsecond') else t
        rb, wb = by(d, 'dram__bytes_read.sum'), by(d, 'dram__bytes_write.sum')
        out.append("%4d %-44s grid %6d x %4d  %9.1f us  tensor %5.1f%%  xu %5.1f%%  dram %5.1f%%  rd %8.2f MB  wr %8.2f MB  (%6.0f GB/s)  L2hit %5.1f%%" % (
            i, d['name'][:44], d['launch__grid_size'], d['launch__block_size'], t_us,
            d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'], d['sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'],
            d['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'], rb / 1e6, wb / 1e6, (rb + wb) / t_us / 1e3, d['lts__t_sector_hit_rate.pct']))
    open(dst, 'w').write("\n".join(out) + "\n")
    print("\n".join(out))

main()
