"""CPU tier for the N>1 path: world_size-2 gloo run of the job plumbing bench.py uses (barrier, max/sum
reduction of per-rank scalars) and of the channel-slab rule shared with the C library's sharding."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    from dsp_b200.dist import Job, channel_slab
    job = Job(backend="gloo")
    job.barrier()
    mx = job.reduce_max(10.0 + rank)           # per-rank "elapsed time": the job time is the slowest rank
    sm = job.reduce_sum(100.0 * (rank + 1))    # per-rank launch counts add up
    b, e = channel_slab(2048, world, rank)
    n = job.reduce_sum(e - b)
    job.barrier()
    job.close()
    q.put((rank, mx, sm, n, b, e))


def test_two_rank_gloo_job():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [11.0, 11.0]
    assert [r[2] for r in res] == [300.0, 300.0]
    assert [r[3] for r in res] == [2048.0, 2048.0]
    assert (res[0][4], res[0][5], res[1][4], res[1][5]) == (0, 1024, 1024, 2048)


@pytest.mark.parametrize("total,parts", [(256, 4), (7, 3), (2048, 8), (5, 5), (3, 2)])
def test_channel_slabs_partition_exactly(total, parts):
    sys.path.insert(0, ROOT)
    from dsp_b200.dist import channel_slab
    covered = []
    for i in range(parts):
        b, e = channel_slab(total, parts, i)
        assert e >= b
        covered += list(range(b, e))
    assert covered == list(range(total))
