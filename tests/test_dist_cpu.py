"""Multi-process host logic on the CPU (gloo, world_size 2): every rank materialises only its own views, one all-gather
makes all segment lists resident, the pair shards are disjoint and complete.  Mirrors what bench.py does over NCCL."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from line3dpp_b200 import shard, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from line3dpp_b200 import shard, synth
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    V, N = 8, 120
    lo, hi = shard.view_range(rank, world, V)
    sc = synth.make_scene_views(V, N, 5, "ring2", range(lo, hi))
    mine = torch.from_numpy(np.concatenate([sc.segs[v] for v in range(lo, hi)]))
    assert mine.shape == ((hi - lo) * N, 4)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)                       # the ONE collective of the data path
    allsegs = torch.cat(parts).numpy()
    full = synth.make_scene(V, N, 5, "ring2")
    assert np.array_equal(allsegs, np.concatenate(full.segs)), "all-gathered segment lists differ from the single-process scene"
    pairs = synth.view_pairs(sc.neighbors)
    my_pairs = shard.rank_pairs(pairs, rank, world, V)
    cnt = torch.tensor([len(my_pairs)]); tot = cnt.clone(); dist.all_reduce(tot)
    assert tot.item() == len(pairs)
    np.save(os.path.join(%r, f"pairs_{rank}.npy"), my_pairs)
    dist.barrier(); dist.destroy_process_group()
    print("rank", rank, "ok", len(my_pairs))
""")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_gloo_allgather_and_pair_shards(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % (ROOT, str(tmp_path)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    a, b = np.load(tmp_path / "pairs_0.npy"), np.load(tmp_path / "pairs_1.npy")
    allp = synth.view_pairs(synth.ring_neighbors(8, 2))
    got = sorted(map(tuple, np.concatenate([a, b]).tolist()))
    assert got == sorted(map(tuple, allp.tolist())) and len(set(got)) == len(got)


def test_shard_helpers():
    pairs = synth.view_pairs(synth.ring_neighbors(1000 * 4, 5))
    seen = 0
    for r in range(4):
        p = shard.rank_pairs(pairs, r, 4, 4000)
        lo, hi = shard.view_range(r, 4, 4000)
        assert ((p[:, 0] >= lo) & (p[:, 0] < hi)).all()
        assert 4900 <= len(p) <= 5100          # balanced: ~5000 view pairs per GPU (BASELINE.json configs[3] per GPU)
        seen += len(p)
    assert seen == len(pairs) == 20000


def test_balanced_split_is_contiguous_complete_and_balanced():
    from line3dpp_b200 import dist as l3dist
    rng = np.random.default_rng(5)
    for n, parts in ((0, 3), (1, 4), (7, 7), (1000, 8), (5000, 3)):
        cost = rng.integers(0, 9_000_000, n)
        b = l3dist.balanced_split(cost, parts)
        assert b[0] == 0 and b[-1] == n and (np.diff(b) >= 0).all()
        if n >= 1000:
            share = np.array([cost[b[r]:b[r + 1]].sum() for r in range(parts)], float)
            assert share.max() <= share.mean() + cost.max()          # never worse than one item off the ideal
    b = l3dist.balanced_split(np.zeros(10, np.int64), 4)                # all-empty views: split by count
    assert b[0] == 0 and b[-1] == 10 and (np.diff(b) >= 2).all()
    b = l3dist.balanced_split([5, 0, 0, 0, 5], 2)
    assert list(b) in ([0, 1, 5], [0, 2, 5], [0, 3, 5], [0, 4, 5])


EXCHANGE_WORKER = textwrap.dedent("""
    import os, sys, numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from line3dpp_b200 import dist as l3dist
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    rows, knn = 1000, 3
    rng = np.random.default_rng(11)
    full_counts = rng.integers(0, knn + 1, rows).astype(np.int32)
    full_recs = rng.integers(0, 255, (rows * knn, l3dist.REC_BYTES)).astype(np.uint8)
    bounds = [0, 0, rows] if %r else [0, 377, rows]              # an empty share must be fine
    counts, recs = np.zeros_like(full_counts), np.zeros_like(full_recs)
    a, b = bounds[rank], bounds[rank + 1]
    counts[a:b] = full_counts[a:b]; recs[a * knn:b * knn] = full_recs[a * knn:b * knn]
    ct, rt = torch.from_numpy(counts.view(np.uint8)), torch.from_numpy(recs.reshape(-1))
    l3dist.exchange_rows(ct, rt, bounds, knn)
    assert np.array_equal(counts, full_counts) and np.array_equal(recs, full_recs), "rows missing after the exchange"
    dist.barrier(); dist.destroy_process_group()
    print("rank", rank, "exchange ok")
""")


def test_two_rank_gloo_match_row_exchange(tmp_path):
    for empty_share in (False, True):
        script = tmp_path / f"xworker_{int(empty_share)}.py"
        script.write_text(EXCHANGE_WORKER % (ROOT, empty_share))
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
               "--master-port", str(_free_port()), str(script)]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
