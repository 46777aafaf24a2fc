"""Summarise ncu outputs brought back in gpurun_out/ into profiles/ (tracked).
usage: summarize_ncu.py <tag> [launches.csv] [report.ncu-rep ...]"""
import collections, csv, re, subprocess, sys, os

tag = sys.argv[1]
out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
os.makedirs(out_dir, exist_ok=True)
for path in sys.argv[2:]:
    if path.endswith(".csv"):
        lines = [l for l in open(path) if not l.startswith("==")]
        agg = collections.defaultdict(lambda: [0, 0.0])
        for row in csv.DictReader(lines):
            if row.get("Metric Name") != "gpu__time_duration.sum":
                continue
            name = re.sub(r"\(.*", "", row["Kernel Name"])
            v = float(row["Metric Value"].replace(",", ""))
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
            agg[name][0] += 1
            agg[name][1] += v
        tot = sum(v[1] for v in agg.values())
        with open(os.path.join(out_dir, tag + "_launches_summary.txt"), "w") as f:
            f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
            f.write("# source: %s ; total %.2f ms over %d launches\n" % (os.path.basename(path), tot, sum(v[0] for v in agg.values())))
            for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write("%-60s n=%5d %10.3f ms %6.2f%%\n" % (k[:60], v[0], v[1], 100 * v[1] / tot))
        print(open(os.path.join(out_dir, tag + "_launches_summary.txt")).read())
    else:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        keep = re.compile(r"Kernel Name|gpu__time_duration.sum|dram__bytes_(read|write).sum$|dram__throughput.avg.pct|gpu__dram_throughput|"
                          r"dram__cycles_active.avg.pct|sm__warps_active.avg.pct|launch__registers_per_thread|launch__grid_size|launch__block_size|"
                          r"launch__shared_mem_per_block_dynamic|smsp__inst_executed.sum$|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|"
                          r"sm__throughput.avg.pct|lts__t_bytes.sum$|lts__t_sector_hit_rate.pct|l1tex__t_sector_hit_rate.pct|sm__pipe_tensor|"
                          r"smsp__average_warp.*issue_stalled.*_per_warp_active.pct|sm__inst_executed_pipe_fp64|sm__cycles_elapsed.avg$")
        name = os.path.splitext(os.path.basename(path))[0]
        with open(os.path.join(out_dir, "%s_%s_raw.txt" % (tag, name)), "w") as f:
            f.write("# ncu --set full --clock-control none --import-source on ; selected raw metrics per captured launch\n")
            for r in rows[2:]:
                f.write("----\n")
                for i, h in enumerate(hdr):
                    if keep.search(h):
                        f.write("%-95s %s %s\n" % (h, r[i], units[i]))
        print("wrote", name)
