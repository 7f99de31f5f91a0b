"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches.txt
    python profiles/summarize.py full     gpurun_out/prof_gemm_r1.ncu-rep > profiles/r1_tc_gemm_full.txt
"""
import collections
import csv
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
           "sm__cycles_active.avg", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
           "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return re.sub(r"nf::\(anonymous namespace\)::|nf::|void |unnamed>::", "", name).strip()


def launches(path):
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    agg, tot, n = collections.OrderedDict(), 0.0, 0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        unit = row["Metric Unit"]
        ms = v / 1e6 if unit.startswith("n") else (v / 1e3 if unit.startswith("u") else v)
        a = agg.setdefault(short(row["Kernel Name"]), [0, 0.0])
        a[0] += 1; a[1] += ms; tot += ms; n += 1
    print("# ncu --metrics gpu__time_duration.sum --clock-control none  (cold-cache, serialised: compare SHARES)")
    print("# %d launches, %.3f ms total" % (n, tot))
    print("%-48s %6s %10s %7s %9s" % ("kernel", "count", "total_ms", "share", "avg_ms"))
    for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-48s %6d %10.3f %6.1f%% %9.3f" % (k[:48], c, ms, 100 * ms / tot, ms / c))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# ncu --set full --clock-control none --import-source on   (%s)" % path)
    for r in rows[2:]:
        print("---- %s" % short(r[idx["Kernel Name"]]))
        for m in METRICS:
            if m in idx:
                print("  %-72s %s %s" % (m, r[idx[m]], units[idx[m]]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
