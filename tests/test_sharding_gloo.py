"""The N > 1 host logic on CPU: world_size-2 `gloo` processes exercise the patient sharding of batch prediction
and the round-robin chain deal + gather.  No GPU: the per-rank scorer is a deterministic stand-in."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from mogp_b200.parallel import shard_bounds, chains_for_rank

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_balance():
    for n in (1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            if n < world:
                continue
            edges = [shard_bounds(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def test_chain_deal():
    deal = [chains_for_rank(64, 8, r) for r in range(8)]
    assert sorted(sum(deal, [])) == list(range(64))
    assert all(len(d) == 8 for d in deal)
    assert chains_for_rank(5, 2, 1) == [1, 3]


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np
    import torch.distributed as dist
    from mogp_b200 import parallel
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    rs = np.random.RandomState(0)
    P, T, K = 37, 6, 5
    X = rs.uniform(0, 8, (P, T)); Y = rs.randn(P, T); a = rs.randn(P)
    def scorer(Xs, Ys, As):   # deterministic stand-in for MoGP_constrained.predict_batch
        s = np.stack([np.sin(Xs).sum(1) * (k + 1) + Ys.sum(1) + As for k in range(K + 1)], axis=1)
        e = np.exp(s - s.max(1, keepdims=True))
        return e / e.sum(1, keepdims=True)
    full = parallel.predict_sharded(scorer, X, Y, a)
    ref = scorer(X, Y, a)
    assert full.shape == ref.shape and np.array_equal(full, ref), "sharded prediction differs"
    # fewer patients than ranks: the rank with an empty shard contributes a 0-row block instead of raising / hanging
    one = parallel.predict_sharded(scorer, X[:1], Y[:1], a[:1])
    assert one.shape == (1, K + 1) and np.array_equal(one, scorer(X[:1], Y[:1], a[:1])), "one-patient sharded prediction differs"
    mine = {{c: dict(z=np.full(4, c, dtype=np.int32), ll=np.array([c * 0.5])) for c in parallel.chains_for_rank(7, world, rank)}}
    allc = parallel.gather_chain_results(mine, 7)
    assert [int(r['z'][0]) for r in allc] == list(range(7))
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_world_size_two_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, out
        assert "rank {} ok".format(rank) in out
