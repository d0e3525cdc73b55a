"""Count the SASS mnemonics that prove tcgen05 / tensor memory / TMA use, per kernel of libspnb.so.

    python tools/sass_mnemonics.py [libspnb.so] [kernel-name-substring ...]

UTCHMMA = tcgen05.mma (kind::tf32 / f16), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA 1-D
bulk copy), SYNCS = mbarrier try_wait / arrive / expect_tx, LDGSTS = cp.async, RED / ATOMG = global float atomics.
"""
import collections
import re
import subprocess
import sys

WANT = ("UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UTMALDG", "SYNCS", "LDGSTS",
        "RED", "ATOMG", "FFMA", "HMMA", "LDS", "STS")


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else "smoothparticlenets_b200/libspnb.so"
    subs = sys.argv[2:]
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    names = {}
    cur = None
    counts = collections.OrderedDict()
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            for w in WANT:
                if op == w or op.startswith(w + "."):
                    counts[cur][w] += 1
    dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.split("\n")
    for mangled, d in zip(counts, dem):
        names[mangled] = re.sub(r"\(.*", "", d.replace("(anonymous namespace)", "{anon}"))
    print("# %s" % " ".join(sys.argv))
    print("# " + __doc__.strip().split("\n\n")[-1].replace("\n", "\n# "))
    for k, c in counts.items():
        n = names[k]
        if subs and not any(s in n for s in subs):
            continue
        if not subs and not any(c[w] for w in WANT[:10]):
            continue
        print("%-100s %s" % (n[:100], "  ".join("%s %d" % (w, c[w]) for w in WANT if c[w])))


if __name__ == "__main__":
    main()
