"""Static SASS statistics per kernel of a .o / .so (no GPU needed): instruction count, opcode mix."""
import re
import subprocess
import sys
from collections import Counter

txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
pat = sys.argv[2] if len(sys.argv) > 2 else ""
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n")[0]
    if pat and pat not in name:
        continue
    ins = re.findall(r"^\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?(\S+)", f, flags=re.M)
    c = Counter(i.split(".")[0] for i in ins)
    short = re.sub(r"_ZN\d+_GLOBAL__N__[0-9a-f_]+cu_[0-9a-f]{8}", "", name)[:48]
    keys = ("MUFU", "FFMA", "FMUL", "FADD", "CALL", "BRA", "BSSY", "LDL", "STL", "LDC", "LDCU")
    print(f"{short:48s} n={len(ins):5d} " + " ".join(f"{k}={c[k]}" for k in keys))
