"""Summaries of ncu reports exported as CSV:
  ncu -i X.ncu-rep --page raw --csv > X_raw.csv ;  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X_src.csv
  python tools/ncu_summary.py raw X_raw.csv | src X_src.csv [top]"""
import csv
import re
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "smsp__sass_average_branch_targets_threads_uniform.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("----")
        for w in WANT:
            if w in idx:
                print(w, "=", r[idx[w]], units[idx[w]])
        st = [(hdr[i], r[i]) for i in range(len(hdr))
              if re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio", hdr[i])]
        st = sorted(st, key=lambda x: -float(x[1] or 0))[:6]
        print("stalls:", [(a.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), b)
                          for a, b in st])


def src(path, top=40, only=None):
    """per kernel: source lines by warp-stall samples (where the warps wait) with their share of
    the executed warp instructions"""
    csv.field_size_limit(10 ** 9)
    rows = list(csv.reader(open(path)))
    cur, fn, hdr = None, None, None
    agg = {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1]
            continue
        if len(r) == 2 and r[0] == "Function Name":
            fn = re.sub(r"\(.*", "", r[1]).replace("rjb::", "").replace("void ", "")
            continue
        if len(r) > 3 and r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            try:
                ln = int(r[0])
            except ValueError:
                continue
            ie = r[hdr.index("Instructions Executed")]
            if ie == "":
                continue
            te = r[hdr.index("Thread Instructions Executed")] if "Thread Instructions Executed" in hdr else "0"
            ss = r[hdr.index("Warp Stall Sampling (All Samples)")] if "Warp Stall Sampling (All Samples)" in hdr else "0"
            a = agg.setdefault(fn, {}).setdefault((cur.split("/")[-1], ln, r[1].strip()[:100]), [0, 0, 0])
            a[0] += int(ie)
            a[1] += int(te or 0)
            a[2] += int(ss or 0)
    for fn, d in agg.items():
        if only and only not in fn:
            continue
        tot = sum(v[0] for v in d.values()) or 1
        tots = sum(v[2] for v in d.values()) or 1
        print("== %s: %d warp instructions, %d stall samples" % (fn, tot, tots))
        for k, v in sorted(d.items(), key=lambda kv: -kv[1][2])[:top]:
            print("  stall %5.1f%%  inst %5.1f%%  thr/inst %5.1f  %s:%d  %s"
                  % (100 * v[2] / tots, 100 * v[0] / tot, v[1] / max(1, v[0]), k[0], k[1], k[2]))


if __name__ == "__main__":
    if sys.argv[1] == "raw":
        raw(sys.argv[2])
    else:
        src(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40, sys.argv[4] if len(sys.argv) > 4 else None)
