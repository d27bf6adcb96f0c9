"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA (tcgen05.mma),
LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA), UBLKCP (bulk copies), plus the legacy HMMA (mma.sync).
    python profiles/sass_summary.py > profiles/sass_summary.md      (needs cuobjdump; runs on the CPU-only build box)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "plbert_b200", "libplbert_b200.so")
PATS = [("UTCHMMA", r"\bUTCHMMA"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"),
        ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"), ("STSM", r"\bSTSM"), ("REDG", r"\bREDG"), ("FFMA2/FMUL2/FADD2", r"\bF(FMA|MUL|ADD)2\b"),
        ("MUFU", r"\bMUFU"), ("HMMA (mma.sync)", r"\bHMMA")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"^void ", "", name)
            name = re.sub(r"plb::\(anonymous namespace\)::|\(anonymous namespace\)::", "", name)
            cur = re.sub(r"\(.*", "", name)
            order.append(cur)
            continue
        if cur is None:
            continue
        for label, pat in PATS:
            if re.search(pat, line):
                counts[cur][label] += 1
    print("# SASS summary of `plbert_b200/libplbert_b200.so` (sm_100a) — `python profiles/sass_summary.py`\n")
    print("Counts of instructions per kernel in the shipped library (cuobjdump -sass).  `UTCHMMA` = tcgen05.mma, `LDTM` / `STTM` ="
          " tcgen05.ld / st, `UTMALDG` / `UTMASTG` = TMA tensor loads / stores, `UBLKCP` = cp.async.bulk, `STSM` = stmatrix.\n")
    labels = [l for l, _ in PATS]
    print("| kernel | " + " | ".join(labels) + " |")
    print("|---|" + "---:|" * len(labels))
    for k in order:
        c = counts[k]
        print(f"| `{k}` | " + " | ".join(str(c[l]) if c[l] else "" for l in labels) + " |")


if __name__ == "__main__":
    sys.exit(main())
