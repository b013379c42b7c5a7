"""Turn an ncu launch list (--metrics gpu__time_duration.sum --csv) and a full capture (.ncu-rep) into the tables kept
under profiles/.  Usage: python tools/summarize_ncu.py launches.csv prof.ncu-rep > profiles/rNN_summary_tables.md"""
import csv, collections, subprocess, sys, io

def launches(path):
    rows = [r for r in csv.reader(open(path, newline="")) if r]
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]; ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= iv: continue
        v = float(r[iv].replace(",", "")); u = r[iu]
        us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
        a = agg.setdefault(r[ik], [0, 0.0]); a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|")
    for k, (n, us) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("| `%s` | %d | %.3f | %.1f | %.3f |" % (k[:70], n, us / 1e3, us / n, us / tot))

def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]; idx = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__icc_request_hit_rate.pct", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
            "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size"]
    names = [r[idx["Kernel Name"]][:28] for r in rows[2:]]
    print("| metric | " + " | ".join("`%s`" % n for n in names) + " |\n|---|" + "---|" * len(names))
    for w in want:
        if w in idx:
            print("| %s (%s) | " % (w, units[idx[w]]) + " | ".join(r[idx[w]] for r in rows[2:]) + " |")

if __name__ == "__main__":
    launches(sys.argv[1]); print(); full(sys.argv[2])
