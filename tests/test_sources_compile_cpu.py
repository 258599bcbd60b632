"""Every Python source of the repo byte-compiles (a syntax error in a torchrun-only tool or in bench.py must not wait for a GPU box to show)."""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_python_source_compiles():
    bad, paths = [], []
    for top in ("aerial_gym_simulator_b200", "tools", "oracle", "tests"):
        for dp, dn, fn in os.walk(os.path.join(ROOT, top)):
            dn[:] = [d for d in dn if d not in ("__pycache__", "_build", "build", "resources")]
            for f in fn:
                if f.endswith(".py"):
                    paths.append(os.path.join(dp, f))
    for path in paths + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]:
        try:
            compile(open(path, encoding="utf-8").read(), path, "exec")
        except SyntaxError as e:
            bad.append(f"{path}: {e}")
    assert not bad, "\n".join(bad)
