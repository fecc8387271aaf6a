"""SASS evidence of what the shipped library executes: per kernel, counts of the tcgen05 / TMA / TMEM mnemonics
(UTCHMMA = tcgen05.mma, UTMALDG = TMA load, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit), of legacy
HMMA (must be absent) and of global atomics / reductions (ATOMG / RED.: only the GroupNorm grid-barrier counter).
    python tests/sass_mnemonics.py > profiles/rNN_sass_mnemonics.json"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "stable-fast_b200", "sfast_b200", "libsfb200.so")
WANT = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "HMMA", "MUFU.EX2", "ATOMG", "RED."]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    cur, per, tot = None, collections.defaultdict(collections.Counter), collections.Counter()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(2)
            for w in WANT:
                if op.startswith(w) or (w == "RED." and op == "RED"):
                    per[cur][w] += 1
                    tot[w] += 1

    def short(n):
        r = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
        return re.sub(r"\(.*", "", r)[:100]
    json.dump({"library": os.path.relpath(SO, ROOT), "totals": dict(tot),
               "per_kernel": {short(k): dict(v) for k, v in per.items()}}, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
