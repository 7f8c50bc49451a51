"""N > 1 host logic on CPU: two gloo ranks shard a batch, run their shard (through the oracle here -- no GPU),
and the gathered result equals the single-process result; the bench's max-over-ranks timing reduce works."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cases
from simple_image_compression_network_b200.shard import shard_range, weak_range


def test_shard_ranges_cover_the_batch():
    for n in (1, 7, 8, 4096, 4099):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert weak_range(4096, 3) == (3 * 4096, 4 * 4096)
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_img, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    oracle.set_threads(1)
    d = cases.CASES["c2d_b"]
    inp = cases.make_inputs(d, num_reps=n_img)
    s = oracle.query(d)
    b, e = shard_range(n_img, rank, world)
    mine = oracle.run_layer(d, inp["in_words"][b * s.in_bytes_per_image:e * s.in_bytes_per_image], inp["weights"], None, inp["bias"],
                            num_reps=e - b)
    # no data-path collective: results stay on their rank; only the test gathers them to compare
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([mine.size], dtype=torch.int64))
    cap = max(int(x.item()) for x in sizes)  # gloo all_gather wants equal sizes: pad the ragged shard
    padded = torch.zeros(cap, dtype=torch.uint8)
    padded[: mine.size] = torch.from_numpy(mine)
    bufs = [torch.zeros(cap, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(bufs, padded)
    bufs = [b[: int(n.item())] for b, n in zip(bufs, sizes)]
    # the bench's timing rule: barrier, then MAX over ranks of the per-rank elapsed time
    dist.barrier()
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        full = oracle.run_layer(d, inp["in_words"], inp["weights"], None, inp["bias"], num_reps=n_img)
        q.put((bool(np.array_equal(np.concatenate([x.numpy() for x in bufs]), full)), float(t.item())))
    dist.destroy_process_group()


def test_two_rank_batch_sharding_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok, "sharded result differs from the single-process result"
    assert tmax == 11.0


def test_library_split_is_the_host_split(fcb_lib):
    """fcb_shard_range (the split fcb_pool_run applies over its replicas) == shard.shard_range, and it tiles the batch."""
    import ctypes
    for n in (0, 1, 7, 8, 9, 4096, 4099):
        for w in (1, 2, 3, 8):
            prev = 0
            for r in range(w):
                b, e = ctypes.c_uint32(), ctypes.c_uint32()
                assert fcb_lib.fcb_shard_range(n, r, w, ctypes.byref(b), ctypes.byref(e)) == 0
                assert (b.value, e.value) == shard_range(n, r, w)
                assert b.value == prev
                prev = e.value
            assert prev == n
    b, e = ctypes.c_uint32(), ctypes.c_uint32()
    assert fcb_lib.fcb_shard_range(4, 2, 2, ctypes.byref(b), ctypes.byref(e)) == -1
