"""Count the SASS mnemonics that prove the sm_100a paths (tcgen05: UTCHMMA / LDTM / UTCBAR; TMA: UTMALDG / UBLKCP; mbarrier: SYNCS; warp
reductions: REDUX / VOTE / SHFL; fp64: DFMA / DMUL / DADD) per kernel of the built library.  Usage: python profiles/sass_mnemonics.py > out.tsv"""
import collections
import re
import subprocess
import sys
from pathlib import Path

KEYS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "REDUX", "VOTE", "SHFL", "MUFU", "HMMA", "DFMA", "DMUL",
        "DADD", "ATOMS", "ATOMG"]


def main() -> None:
    so = Path(__file__).resolve().parents[1] / "tiatoolbox_b200" / "libtiab200.so"
    sass = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True, check=True).stdout
    per, fn = collections.defaultdict(collections.Counter), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        if fn is None or "/*" not in line:
            continue
        for k in KEYS:
            if re.search(r"\b" + re.escape(k) + r"\b|\b" + re.escape(k) + r"\.", line):
                per[fn][k] += 1
    names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True, check=True).stdout.splitlines()
    total = collections.Counter()
    print("kernel\t" + "\t".join(KEYS))
    for raw, name in sorted(zip(per, names), key=lambda t: t[1]):
        total.update(per[raw])
        short = re.sub(r"\(.*", "", name)
        print(short + "\t" + "\t".join(str(per[raw].get(k, 0)) for k in KEYS))
    print("TOTAL (%d kernels)\t" % len(per) + "\t".join(str(total.get(k, 0)) for k in KEYS))
    sys.stderr.write(f"{len(per)} kernels\n")


if __name__ == "__main__":
    main()
