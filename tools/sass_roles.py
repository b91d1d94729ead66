"""Which SASS instructions of cqa_kernel spill, and in which register-budget region (setmaxnreg) they sit."""
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "tvqaplus_b200/libstage_b200.so"
pat = sys.argv[2] if len(sys.argv) > 2 else "cqa_kernelILb0"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    name = part.split("\n", 1)[0]
    if pat not in name:
        continue
    lines = part.split("\n")
    marks = [i for i, l in enumerate(lines) if "USETMAXREG" in l]
    print(name, len(lines), "lines; setmaxnreg at", marks)
    for i, l in enumerate(lines):
        if re.search(r"\b(STL|LDL)\b", l):
            print(i, l.strip()[:100])
    print("UTCHMMA", sum("UTCHMMA" in l for l in lines), "BRA.U.ANY", sum("BRA.U.ANY" in l for l in lines))
