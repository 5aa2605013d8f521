"""CPU test of the N>1 host logic (world_size 2, gloo): contiguous utterance shards and the single end-of-run
all-gather of per-utterance statistics."""
import os
import socket

import numpy as np
import pytest

from distant_speech_recognition_b200 import sharding


def test_shard_range_partition():
    for total in (0, 1, 7, 256, 1000, 8192):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = sharding.shard_range(total, world, rank)
    # fake per-utterance statistics that encode the global utterance index
    local = np.stack([np.arange(a, b) * 10.0, np.full(b - a, 317.0), np.arange(a, b) % 5], axis=1) if b > a else np.zeros((0, 3))
    allst = sharding.gather_stats(local)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, allst))


@pytest.mark.parametrize("total", [7, 8])
def test_gather_stats_world2_gloo(total):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in (0, 1):
        st = res[r]
        assert st.shape == (total, 3)
        assert np.array_equal(st[:, 0], np.arange(total) * 10.0)   # rank order == utterance order, ragged shard trimmed
        assert np.array_equal(st[:, 2], np.arange(total) % 5)
    s = sharding.summarize(res[0], 256)
    assert s["utterances"] == total and s["frames"] == 317.0 * total


def test_gather_stats_single_process_passthrough():
    st = np.arange(12, dtype=np.float64).reshape(4, 3)
    assert np.array_equal(sharding.gather_stats(st), st)
