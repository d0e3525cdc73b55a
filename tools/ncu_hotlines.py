"""Per-source-line instruction counts and stall samples of one profiled kernel.

    python tools/ncu_hotlines.py REPORT.ncu-rep CUBIN KERNEL_SUBSTRING SOURCE.cu [N]

ncu's SASS page gives executed instructions / samples per address; nvdisasm -g gives the source line of every
address of the same cubin (extract with `cuobjdump -xelf all libspnb.so`).  Joined here by instruction order.
"""
import csv
import io
import re
import subprocess
import sys


def main():
    rep, cubin, kern, srcfile = sys.argv[1:5]
    n = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = rows[1]
    ie, sm = h.index("Instructions Executed"), h.index("# Samples")
    body = [(r[h.index("Source")], int(r[ie]), int(r[sm])) for r in rows[2:] if len(r) == len(h)]
    dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.split("\n")
    # the function's section
    start = [i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l][0]
    lines, cur = [], None
    base = srcfile.split("/")[-1]
    for l in dis[start + 1:]:
        if l.startswith("\t.section") or (l.startswith(".text.") and kern not in l):
            break
        m = re.search(r'//## File "(.*?)", line (\d+)', l)
        if m:
            f = m.group(1).split("/")[-1]
            cur = int(m.group(2)) if f == base else "%s:%s" % (f, m.group(2))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            lines.append(cur)
    assert abs(len(lines) - len(body)) <= 2, (len(lines), len(body))
    tot = sum(b[1] for b in body) or 1
    tots = sum(b[2] for b in body) or 1
    agg = {}
    for (s, ni, ns), ln in zip(body, lines):
        d = agg.setdefault(ln, [0, 0])
        d[0] += ni
        d[1] += ns
    src = open(srcfile).read().split("\n")
    print("total warp instructions %d, samples %d" % (tot, tots))
    for ln, (ni, ns) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
        text = src[ln - 1].strip()[:105] if isinstance(ln, int) else ""
        print("%6.2f%% inst %5.1f%% smp  %-22s %s" % (100.0 * ni / tot, 100.0 * ns / tots, ln, text))


if __name__ == "__main__":
    main()
