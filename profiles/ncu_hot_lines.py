"""Summarise `ncu -i <rep> --page source --csv --print-source cuda,sass --kernel-name regex:<k> --launch-count 1`:
stall samples and executed warp instructions per source line.  Usage: python profiles/ncu_hot_lines.py src.csv [top]"""
import csv, sys


def summarize(path, top=30):
    rows = list(csv.reader(open(path, newline="")))
    out, cur = [], None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] not in ("Function Name", "Line No") and r[0] and len(r) > 7 and r[2] == "-":
            try:
                out.append((cur, int(r[0]), r[1].strip(), int(r[4]), int(r[7])))
            except ValueError:
                pass
    tot = sum(o[3] for o in out) or 1
    toti = sum(o[4] for o in out) or 1
    print("total stall samples %d, executed warp instructions %d" % (tot, toti))
    print("%8s %8s  %s" % ("samples", "instr", "source line"))
    for f, l, src, s, i in sorted(out, key=lambda o: -o[3])[:top]:
        print("%7.1f%% %7.1f%%  %s:%d  %s" % (100 * s / tot, 100 * i / toti, f, l, src[:110]))


if __name__ == "__main__":
    summarize(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
