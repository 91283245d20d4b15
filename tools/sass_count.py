"""Static SASS opcode counts of the coarse kernel in one or more objects / libraries (cuobjdump -sass), e.g. to compare the
activation forms of tuning builds:

    python tools/sass_count.py rails_b200/build/mol_coarse_sm100.o rails_b200/lib/libmol_b200_h2_3e.so

E2 is inlined once and E3 four times in mol_coarse_kernel<8, 32>, so e.g. MUFU.TANH = 2 * (E2 pairs on the MUFU) + 4 * 64.
"""
import collections
import re
import subprocess
import sys

KEYS = ["MUFU.TANH", "MUFU.EX2", "MUFU.RCP", "HFMA2", "HFMA2.RELU", "HMUL2", "HADD2", "FFMA2", "FMUL2", "FADD2", "FFMA",
        "FMNMX", "F2FP.F16", "LOP3", "MOV", "PRMT", "UTCHMMA", "LDTM", "STTM", "LDL", "STL"]


def counts(path, kernel="mol_coarse_kernelILi8ELi32"):
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    out, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            parts = m.group(1).split(".")
            key = parts[0]
            if parts[0] in ("MUFU", "F2FP") and len(parts) > 1:
                key = ".".join(parts[:2])
            if parts[0] == "HFMA2" and "RELU" in parts:
                key = "HFMA2.RELU"
            out[cur][key] += 1
    return {k: v for k, v in out.items() if kernel in k}


if __name__ == "__main__":
    for path in sys.argv[1:]:
        for name, c in counts(path).items():
            print(path, "total", sum(c.values()), {k: c[k] for k in KEYS if c[k]})
