"""Every Python source of the repo byte-compiles (a syntax error in a torchrun-only tool or in bench.py must not wait for a GPU box to show)."""
import os
import py_compile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_python_source_compiles():
    bad = []
    for top in ("aerial_gym_simulator_b200", "tools", "oracle", "tests"):
        for dp, dn, fn in os.walk(os.path.join(ROOT, top)):
            dn[:] = [d for d in dn if d not in ("__pycache__", "_build", "build", "resources")]
            for f in fn:
                if f.endswith(".py"):
                    try:
                        py_compile.compile(os.path.join(dp, f), doraise=True, cfile=os.devnull)
                    except py_compile.PyCompileError as e:
                        bad.append(str(e))
    for f in ("bench.py", "__graft_entry__.py"):
        try:
            py_compile.compile(os.path.join(ROOT, f), doraise=True, cfile=os.devnull)
        except py_compile.PyCompileError as e:
            bad.append(str(e))
    assert not bad, "\n".join(bad)
