"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md, "What proves a
Blackwell-native kernel"): UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG (TMA),
plus registers / shared memory per kernel.  Runs here (no GPU needed):

    python tools/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "quantized_vit_b200", "libqvit_b200.so")
PAT = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "HMMA", "IMMA", "IDP",
       "FFMA2", "MUFU", "SYNCS"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_instr"] += 1
            for p in PAT:
                if op.startswith(p):
                    counts[cur][p] += 1
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    fn = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*SHARED:(\d+)", line)
        if m and fn:
            usage[fn] = (int(m.group(1)), int(m.group(2)))
    dm = demangle(order)
    print(f"# SASS summary of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass / -res-usage; sm_100a)")
    tot = collections.Counter()
    for k in order:
        c = counts[k]
        tot.update(c)
        tags = " ".join(f"{p}={c[p]}" for p in PAT if c[p])
        reg, sh = usage.get(k, (None, None))
        name = re.sub(r"\(.*$", "", dm.get(k, k))
        print(f"{name}\n    instr={c['_instr']} regs={reg} static_smem={sh}  {tags}")
    print("# totals: " + " ".join(f"{p}={tot[p]}" for p in PAT if tot[p]))


if __name__ == "__main__":
    sys.exit(main())
