"""Opcode histogram per kernel of libvaegan_b200.so (cuobjdump -sass): the Blackwell-specific mnemonics the design
rests on (UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA loads / stores, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, PREEXIT / ACQBULK = programmatic dependent launch, LDGMC = multimem.ld_reduce) and what must be
absent (HMMA = legacy mma.sync).   python scripts/sass_histogram.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vae-gan-based-model-for-image-generation-and-denoising_b200", "libvaegan_b200.so")
KEY = ("UTCHMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCCP", "SYNCS", "PREEXIT", "ACQBULK",
       "LDGMC", "HMMA", "ELECT", "REDG", "RED", "ATOMS", "ATOMG", "SHFL", "LDS", "STS", "LDG", "STG")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = per.setdefault(re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0], collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            cur["__total__"] += 1
            op = m.group(1)
            if op in KEY:
                cur[op + (".2CTA" if ".2CTA" in m.group(2) else "")] += 1
    tot = collections.Counter()
    print(f"# {os.path.basename(LIB)}: {len(per)} kernels (sm_100a)\n")
    for name, c in per.items():
        keys = ", ".join(f"{k} {v}" for k, v in sorted(c.items()) if k != "__total__")
        print(f"{name}\n    {c['__total__']} instructions; {keys}")
        tot.update(c)
    print("\n# whole library: " + ", ".join(f"{k} {v}" for k, v in sorted(tot.items()) if k != "__total__"))
    print("# HMMA (legacy mma.sync) present:", "yes" if tot.get("HMMA") else "no")


if __name__ == "__main__":
    sys.exit(main())
