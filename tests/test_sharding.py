"""Multi-GPU host logic on CPU: world-size-2 gloo processes shard a global frame list, each "extracts" its shard (the CPU
oracle stands in for the device here — this is a test of the sharding and gather order, not of the kernels), and the
gathered result must equal the single-process result in input order."""
import os
import socket

import numpy as np
import pytest

from orb_slam2_with_comment_b200 import sharding


def test_shard_ranges_are_contiguous_balanced_and_cover():
    for n in (0, 1, 7, 8, 1024, 8192, 65536 + 3):
        for world in (1, 2, 3, 4, 8):
            r = [sharding.shard_range(n, g, world) for g in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1 and sizes == sharding.shard_sizes(n, world)
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def _worker(rank, world, port, n_frames, q):
    import torch.distributed as dist
    import oracle_lib as ol
    from orb_slam2_with_comment_b200 import synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(n_frames, rank, world)
    ex = ol.Extractor(ol.load_port(), "orbo", 300, 1.2, 8, 20, 7)
    counts, kps, descs = [], [], []
    for f in range(lo, hi):
        kp, d = ex.extract(synth.g_rects(320, 240, 1000 + f))
        counts.append(len(kp)); kps.append(kp); descs.append(d)
    local_counts = np.array(counts, np.int32)
    local_kp = np.concatenate(kps) if kps else np.zeros(0, ol.KP_DTYPE)
    local_desc = np.concatenate(descs) if descs else np.zeros((0, 32), np.uint8)
    all_counts = sharding.gather_ragged(local_counts, dist)
    all_kp = sharding.gather_ragged(local_kp, dist)
    all_desc = sharding.gather_ragged(local_desc, dist)
    # max-over-ranks of a timing scalar, as bench.py does
    import torch
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((all_counts, all_kp.tobytes(), all_desc.tobytes(), float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shards_gather_in_input_order(oracle):
    import torch.multiprocessing as mp
    import oracle_lib as ol
    from orb_slam2_with_comment_b200 import synth
    n_frames, world = 5, 2     # odd on purpose: ragged shards
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    counts, kp_bytes, desc_bytes, tmax = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ex = ol.Extractor(oracle, "orbo", 300, 1.2, 8, 20, 7)
    ref = [ex.extract(synth.g_rects(320, 240, 1000 + f)) for f in range(n_frames)]
    assert counts.tolist() == [len(k) for k, _ in ref]
    assert kp_bytes == np.concatenate([k for k, _ in ref]).tobytes()
    assert desc_bytes == np.concatenate([d for _, d in ref]).tobytes()
    assert tmax == 2.0
    c, kp, desc, off = sharding.merge_frame_results([counts[:2], counts[2:]], [np.zeros((2, 4)), np.zeros((3, 4))],
                                                    [np.zeros((2, 4, 32)), np.zeros((3, 4, 32))])
    assert off[-1] == counts.sum() and len(kp) == 5
