"""Hash the SASS of every translation unit of libaerial_gym_b200.so (comment lines, which carry -lineinfo file/line references,
are dropped).  Used to prove that a source-level refactor done without a GPU leaves the machine code of GPU-verified kernels
untouched (profiles/sass_identity_r1.md):

    python tools/sass_hash.py                 # prints unit, md5, instruction-line count
    python tools/sass_hash.py --save ref.json # ... and stores it;   --check ref.json  compares, exit 1 on any difference"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "aerial_gym_simulator_b200", "build")


def hashes():
    out = {}
    for f in sorted(os.listdir(BUILD)):
        if not f.endswith(".o"):
            continue
        txt = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, f)], capture_output=True, text=True, check=True).stdout
        lines = [l for l in txt.splitlines() if not l.lstrip().startswith("//") and not l.startswith("identifier =")]
        out[f] = {"md5": hashlib.md5("\n".join(lines).encode()).hexdigest(), "lines": len(lines)}
    return out


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    from aerial_gym_simulator_b200 import _build
    _build.build()
    h = hashes()
    for k, v in h.items():
        print(f"{k:24s} {v['md5']}  {v['lines']}")
    if "--save" in sys.argv:
        json.dump(h, open(sys.argv[sys.argv.index("--save") + 1], "w"), indent=1)
    if "--check" in sys.argv:
        ref = json.load(open(sys.argv[sys.argv.index("--check") + 1]))
        bad = [k for k in ref if h.get(k, {}).get("md5") != ref[k]["md5"]]
        print("DIFFERENT:" if bad else "all identical", *bad)
        sys.exit(1 if bad else 0)
