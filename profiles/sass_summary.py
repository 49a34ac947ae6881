"""Opcode histogram per kernel of the built library (`cuobjdump -sass`), the evidence that the hot kernels are
tcgen05 / TMEM / TMA code: UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = cp.async.bulk.tensor,
SYNCS = mbarrier, FFMA2 / FADD2 = packed f32x2, MUFU = special-function unit.  Runs without a GPU.
    python profiles/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "dove_b200" / "libdove_b200.so"
KEEP = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "HMMA", "MUFU",
        "FFMA2", "FADD2", "FMUL2", "FFMA", "FMNMX", "FMNMX3", "F2FP", "IMAD", "LDG", "STG", "LDS", "STS", "BAR", "ELECT",
        "UCGABAR", "ATOM", "RED", "SHFL")


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = kernels.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            op, mods = m.group(1), m.group(2)
            cur[op] += 1
            if op in ("UTCHMMA", "UTMALDG", "LDTM", "STTM", "MUFU"):
                cur[op + mods] += 1
    print(f"# cuobjdump -sass {LIB.name}: instruction counts per kernel (static; selected opcodes)")
    for name, c in kernels.items():
        total = sum(v for k, v in c.items() if "." not in k)
        print(f"\n## {name}   ({total} instructions)")
        sel = [(k, v) for k, v in sorted(c.items()) if k.split(".")[0] in KEEP]
        print("   " + "  ".join(f"{k}={v}" for k, v in sel))


if __name__ == "__main__":
    sys.exit(main())
