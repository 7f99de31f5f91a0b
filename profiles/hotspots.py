"""Aggregate an ncu source page (--page source --csv --print-source sass,cuda) per CUDA source line.

    ncu -i X.ncu-rep --page source --csv --print-source sass,cuda > /tmp/src.csv
    python profiles/hotspots.py /tmp/src.csv [top]
"""
import collections
import csv
import sys


def main(path, top=30):
    rows = list(csv.reader(open(path)))
    starts = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Line No"]
    for si, s in enumerate(starts):
        h = rows[s]
        end = starts[si + 1] - 2 if si + 1 < len(starts) else len(rows)
        fname = rows[s - 1][1] if s > 0 and len(rows[s - 1]) > 1 else "?"
        idx = {n: i for i, n in enumerate(h)}
        i_samp, i_inst = idx["# Samples"], idx["Instructions Executed"]
        stall_cols = [(n, i) for n, i in idx.items() if n.startswith("stall_") and "Not Issued" not in n]
        agg = collections.OrderedDict()
        cur_line, cur_src = None, ""
        tot_s = tot_i = 0
        for r in rows[s + 1:end]:
            if len(r) < len(h):
                continue
            if r[0]:
                cur_line, cur_src = r[0], r[1]
                continue   # a CUDA line row; SASS rows follow with empty line number
            try:
                smp = int(float(r[i_samp] or 0)); ins = int(float(r[i_inst] or 0))
            except ValueError:
                continue
            a = agg.setdefault(cur_line, [cur_src, 0, 0, collections.Counter()])
            a[1] += smp; a[2] += ins; tot_s += smp; tot_i += ins
            for n, i in stall_cols:
                try:
                    v = int(float(r[i] or 0))
                except ValueError:
                    v = 0
                if v:
                    a[3][n] += v
        if tot_s == 0:
            continue
        print("# %s  (%d warp instructions, %d stall samples)" % (fname, tot_i, tot_s))
        print("  inst%  stall%  line  top stall reasons | source")
        for line, (src, smp, ins, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
            reasons = " ".join("%s:%d" % (n[6:], v) for n, v in st.most_common(3))
            print("  %5.2f%% %6.2f%%  L%-4s %-40s | %s" % (100.0 * ins / max(tot_i, 1), 100.0 * smp / tot_s, line, reasons, src.strip()[:110]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
