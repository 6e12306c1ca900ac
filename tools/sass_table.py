"""SASS opcode table of libclskd_sm100.so (`cuobjdump -sass`): per kernel (all template instances summed) the counts of the
opcodes that prove the Blackwell path - profiles/r02_sass_opcodes.md.

    python tools/sass_table.py [out.md]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "speech-enhancement-clskd_b200", "clskd_b200", "libclskd_sm100.so")
OPS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM", "HMMA", "LDSM", "FFMA2", "FADD2", "FMUL2", "LDGSTS"]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_opcodes.md")
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.defaultdict(collections.Counter)
    fn = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            dem = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            dem = dem.replace("(anonymous namespace)::", "")
            fn = re.sub(r"<.*", "", dem.split("(")[0]).split("::")[-1].replace("void ", "")
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
        if m and fn:
            op = m.group(1)
            if op in OPS:
                per[fn][op] += 1
    rows = sorted(per.items(), key=lambda kv: (-kv[1]["UTCHMMA"], -kv[1]["HMMA"], -kv[1]["UTMALDG"], kv[0]))
    with open(out, "w") as f:
        f.write("# SASS opcode table of `libclskd_sm100.so` (round 2; `cuobjdump -sass`, all template instances of a kernel summed; tools/sass_table.py)\n\n")
        f.write("`UTCHMMA` = tcgen05.mma, `UTMALDG` / `UTMASTG` / `UTMAREDG` = TMA load / store / reduce-add, `LDTM` = tcgen05.ld, `UTCBAR` = tcgen05.commit,\n"
                "`HMMA` = mma.sync (LSTM recurrence, first encoder layer, channel Gram), `LDSM` = ldmatrix, `FFMA2` = packed fp32 pairs, `LDGSTS` = cp.async.\n\n")
        f.write("| kernel | " + " | ".join(OPS) + " |\n|---|" + "---:|" * len(OPS) + "\n")
        for name, c in rows:
            f.write("| `%s` | " % name + " | ".join(str(c[o]) if c[o] else "" for o in OPS) + " |\n")
    print(out, len(rows), "kernels")


if __name__ == "__main__":
    main()
