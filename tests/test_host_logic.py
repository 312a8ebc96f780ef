"""CPU tests of the host-side logic: workload generators, sharding (incl. a 2-rank gloo run)."""
import os
import subprocess
import sys

import numpy as np

from sedef_b200 import shard, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_generators_deterministic_and_shaped():
    a = synth.make_pairs_small(50, length=1000, div=0.05)
    b = synth.make_pairs_small(50, length=1000, div=0.05)
    assert np.array_equal(a.q, b.q) and np.array_equal(a.t, b.t) and np.array_equal(a.t_raw, b.t_raw)
    assert (a.qlen == 1000).all() and abs(a.tlen.mean() - 1000) < 15
    assert a.q.max() <= 4 and a.t.max() <= 4
    assert np.array_equal(synth.encode(a.q_raw), a.q) and np.array_equal(synth.encode(a.t_raw), a.t)
    low = (a.q_raw >= ord("a")).mean()
    assert 0.2 < low < 0.8                               # soft-masked
    big = synth.make_pairs_large(3, min_len=2000, max_len=4000)
    assert (big.qlen >= 2000).all() and (big.tlen > 1000).all()
    assert np.array_equal(synth.encode(big.t_raw), big.t)


def test_sedef_matrix():
    m = synth.sedef_matrix().reshape(5, 5)
    assert m[0, 0] == 5 and m[0, 1] == -4 and (m[4] == 0).all() and (m[:, 4] == 0).all()


def test_lpt_partition_balances_and_covers():
    rng = np.random.default_rng(1)
    ql = rng.integers(10, 50000, 5000).astype(np.int32); tl = ql + rng.integers(0, 50, 5000).astype(np.int32)
    work = shard.est_cells(ql, tl, 500)
    for k in (1, 2, 4, 8):
        parts = shard.lpt_partition(work, k)
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(5000))
        loads = np.array([work[p].sum() for p in parts], float)
        assert loads.max() / loads.mean() < 1.01


def test_two_rank_gloo_sharding(tmp_path):
    """world_size-2 gloo run: each rank takes its LPT shard; together they cover every pair exactly once."""
    script = tmp_path / "rank.py"
    script.write_text(
        "import os, sys, numpy as np, torch, torch.distributed as dist\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from sedef_b200 import shard, synth\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "ps = synth.make_pairs_small(301, length=200, div=0.05, len_jitter=80)\n"
        "idx = shard.shard_for_rank(ps.qlen, ps.tlen, 100, r, w)\n"
        "mask = torch.zeros(ps.n, dtype=torch.int32); mask[torch.from_numpy(idx)] = 1\n"
        "dist.all_reduce(mask)\n"
        "work = torch.tensor([float(shard.est_cells(ps.qlen[idx], ps.tlen[idx], 100).sum())])\n"
        "mx = work.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)\n"
        "tot = work.clone(); dist.all_reduce(tot)\n"
        "assert int(mask.min()) == 1 and int(mask.max()) == 1, 'shards must partition the pairs'\n"
        "assert float(mx) / (float(tot) / w) < 1.02, 'LPT shards must be balanced'\n"
        # one marker file per rank: stdout is shared between the ranks and their writes interleave
        f"open(os.path.join({str(tmp_path)!r}, 'rank%d.ok' % r), 'w').write('%d %d' % (r, len(idx)))\n"
        "dist.barrier()\n")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29577", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    sizes = []
    for r in range(2):
        marker = tmp_path / f"rank{r}.ok"
        assert marker.exists(), f"rank {r} did not finish: " + out.stdout + out.stderr
        rr, n = map(int, marker.read_text().split())
        assert rr == r and n > 0
        sizes.append(n)
    assert sum(sizes) == 301
