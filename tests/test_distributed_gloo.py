"""N > 1 host logic on CPU: world_size-2 gloo process groups (127.0.0.1 rendezvous).
Covers the env sharding arithmetic, the observation all-gather (equal and ragged shards) and
the sharding invariance of the device-RNG reset spec (oracle Philox)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aerial_gym_simulator_b200.distributed import ObsAllGather, shard_range


def test_shard_range_partitions_exactly():
    for n, w in [(65536, 8), (10, 3), (7, 8), (1_000_003, 4)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        for (o0, c0), (o1, _) in zip(spans, spans[1:]):
            assert o0 + c0 == o1
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 3, 3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_global, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import hp1_oracle as O
        from oracle import philox

        off, cnt = shard_range(n_global, rank, world)
        # each rank simulates its shard with the oracle; resets keyed by GLOBAL env ids
        model = O.Hp1Model()
        st = O.make_state(model, cnt)
        d = philox.reset_uniforms(11, off + np.arange(cnt), np.zeros(cnt, int), 4)
        t = {k: torch.tensor(v) for k, v in d.items()}
        draws = O.ResetDraws(t["bounds_lo"], t["bounds_hi"], t["state"], None, None, None, None, t["tau_inc"], t["tau_dec"],
                             t["thrust"], t["k_thrust"])
        O.reset_envs(model, st, torch.ones(cnt, dtype=torch.bool), draws)
        g = torch.Generator().manual_seed(0)
        acts = torch.rand(n_global, 4, generator=g) * 2 - 1  # same global action table on every rank
        for _ in range(3):
            obs, *_ = O.position_task_step(model, st, acts[off:off + cnt], torch.zeros(cnt, 3))
        gather = ObsAllGather(cnt, 13, n_global, "cpu")
        full = gather(obs)
        dist.barrier()
        if rank == 0:
            q.put(full.clone().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_global", [64, 37])  # equal and ragged shards
def test_sharded_oracle_equals_single_process(n_global):
    from oracle import hp1_oracle as O
    from oracle import philox

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_global, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process run over all envs
    model = O.Hp1Model()
    st = O.make_state(model, n_global)
    d = philox.reset_uniforms(11, np.arange(n_global), np.zeros(n_global, int), 4)
    t = {k: torch.tensor(v) for k, v in d.items()}
    draws = O.ResetDraws(t["bounds_lo"], t["bounds_hi"], t["state"], None, None, None, None, t["tau_inc"], t["tau_dec"],
                         t["thrust"], t["k_thrust"])
    O.reset_envs(model, st, torch.ones(n_global, dtype=torch.bool), draws)
    g = torch.Generator().manual_seed(0)
    acts = torch.rand(n_global, 4, generator=g) * 2 - 1
    for _ in range(3):
        obs, *_ = O.position_task_step(model, st, acts, torch.zeros(n_global, 3))
    assert np.array_equal(gathered, obs.numpy())  # sharding changes nothing, bit for bit


def test_task_shard_spec_and_registry_args(monkeypatch):
    """host side of the env-sharded task API (task/base_task.py): where world / rank come from, and make_task(..., args=) merging into
    task_config.args without touching the registered defaults' other keys"""
    import types

    from aerial_gym_simulator_b200.task.base_task import BaseTask

    class T(BaseTask):
        reset = reset_idx = step = render = close = lambda self, *a, **k: None

    mk = lambda args: T(types.SimpleNamespace(seed=1, args=args))
    assert mk(None).shard_spec() == (1, 0, False)
    assert mk({"world_size": 4, "rank": 3}).shard_spec() == (4, 3, False)
    assert mk({"world_size": 2, "loopback": True}).shard_spec() == (2, 0, True)
    monkeypatch.setenv("WORLD_SIZE", "8")
    monkeypatch.setenv("RANK", "5")
    assert mk({"shard": "torchrun"}).shard_spec() == (8, 5, False)
    t = mk({"world_size": 1})
    assert t.init_sharding(16, 13, "cpu") is None and t.obs_gather is None  # unsharded: nothing is built
    with pytest.raises(ValueError, match="CUDA"):
        mk({"world_size": 2, "rank": 0}).init_sharding(16, 13, "cpu")

    from aerial_gym_simulator_b200.registry._core import TaskRegistry

    seen = {}

    class Fake:
        def __init__(self, cfg, **kw):
            seen["args"] = dict(cfg.args)

    reg = TaskRegistry()
    reg.register_task("fake", Fake, types.SimpleNamespace(args={"reset_rng": "device"}))
    reg.make_task("fake", args={"world_size": 2, "rank": 1})
    assert seen["args"] == {"reset_rng": "device", "world_size": 2, "rank": 1}
