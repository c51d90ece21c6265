"""Per-source-line instruction / stall-sample shares of one kernel from an ncu report with -lineinfo + --import-source.

    ncu -i rep.ncu-rep --page source --csv --print-source=cuda,sass --kernel-name regex:<k> --launch-count 1 > src.csv
    python tools/ncu_lines.py src.csv [top]
"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    cur_file, hdr, agg, tot_i, tot_s = None, None, collections.OrderedDict(), 0, 0
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if len(r) > 8 and r[0] == "Line No":
            hdr = r
            continue
        if not hdr or len(r) < 9 or not r[0].isdigit():
            continue
        try:
            inst = int(r[7]) if r[7] not in ("-", "") else 0
            samp = int(r[6]) if r[6] not in ("-", "") else 0
        except ValueError:
            continue
        k = (cur_file, int(r[0]))
        a = agg.get(k, (0, 0, r[1].strip()))
        agg[k] = (a[0] + inst, a[1] + samp, a[2])
        tot_i += inst
        tot_s += samp
    print("total warp instructions %d, stall samples %d" % (tot_i, tot_s))
    print("  inst%  samp%  file:line  source")
    for (f, ln), (inst, samp, src) in sorted(agg.items(), key=lambda kv: -kv[1][0] - kv[1][1] * tot_i / max(tot_s, 1))[:top_n]:
        print("%6.1f %6.1f  %s:%d  %s" % (100.0 * inst / max(tot_i, 1), 100.0 * samp / max(tot_s, 1), f, ln, src[:110]))


if __name__ == "__main__":
    main()
