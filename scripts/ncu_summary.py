"""Turn ncu exports into the small text summaries committed under profiles/.
  launches:  python scripts/ncu_summary.py launches gpurun_out/launches_bench.csv
  full:      python scripts/ncu_summary.py full gpurun_out/prof.ncu-rep
  traffic:   python scripts/ncu_summary.py traffic gpurun_out/launches_3d.csv [points_per_launch]
             (csv of --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum)
"""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    iname, ival, iunit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows[1:]:
        v = float(r[ival].replace(",", ""))
        v = v / 1e3 if r[iunit] in ("ns", "nsecond") else v * (1e3 if r[iunit] in ("ms", "msecond") else 1.0)
        name = r[iname].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
    print("# kernel launches of the command (ncu --metrics gpu__time_duration.sum --clock-control none);")
    print("# times are cold-cache / serialised: compare SHARES, not absolutes. total %.3f ms, %d launches" %
          (total / 1e3, len(rows) - 1))
    print("%-70s %8s %12s %8s" % ("kernel", "launches", "total_us", "share"))
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-70s %8d %12.1f %7.2f%%" % (name[:70], n, t, 100 * t / total))


WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
        "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("# ncu --set full --clock-control none --import-source on  (%s)" % path.split("/")[-1])
    for r in rows[2:]:
        print("\n== %s" % r[hdr.index("Kernel Name")])
        for w in WANT:
            if w in hdr:
                print("  %-72s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
        tot = float(r[hdr.index("smsp__pcsamp_sample_count")])
        st = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    st.append((100 * float(r[i]) / tot, h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        print("  warp-state samples: " + ", ".join("%s %.1f%%" % (h, v) for v, h in sorted(st, reverse=True)[:9]))


def traffic(path, npts=None):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    iid, iname, imet = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name")
    ival, iunit = hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3,
             "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = collections.OrderedDict()
    for r in rows[1:]:
        d = per.setdefault(r[iid], {"name": r[iname]})
        d[r[imet]] = float(r[ival].replace(",", "")) * scale.get(r[iunit], 1.0)
    agg = collections.OrderedDict()
    for d in per.values():
        a = agg.setdefault(d["name"].split("(")[0], [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0)
        a[3] += d.get("dram__bytes_write.sum", 0.0)
    print("%-78s %5s %10s %9s %9s %10s" % ("kernel", "n", "avg_us", "rd_GB", "wr_GB", "dram_GB/s"))
    for name, (n, t, rd, wr) in agg.items():
        line = "%-78s %5d %10.1f %9.3f %9.3f %10.0f" % (name[:78], n, t / n, rd / n / 1e9, wr / n / 1e9,
                                                       (rd + wr) / (t * 1e-6) / 1e9 if t else 0.0)
        if npts and n and (rd + wr) / n > 1e9:
            line += "  %.1f B/pt moved" % ((rd + wr) / n / npts)
        print(line)


if __name__ == "__main__":
    fn = {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]]
    if sys.argv[1] == "traffic" and len(sys.argv) > 3:
        fn(sys.argv[2], float(sys.argv[3]))
    else:
        fn(sys.argv[2])
