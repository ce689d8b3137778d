"""CPU, world_size 2 over gloo: picture-range sharding and the ordered gather of per-picture byte buffers."""
import os
import socket

import torch.multiprocessing as mp

from wrenc_b200.sharding import shard_range


def test_shard_range_partitions_contiguously():
    for n in (0, 1, 7, 240, 241):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from wrenc_b200.sharding import gather_in_order, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 7
    b, e = shard_range(n, world, rank)
    bufs = [bytes([p]) * (p * 37 % 11 + (0 if p == 3 else 1)) for p in range(b, e)]  # variable lengths, picture 3 empty
    out = gather_in_order(bufs, dst=0)
    from wrenc_b200.sharding import gather_in_order_host
    for _ in range(2):  # the host shared-memory variant, twice (per-call sequence numbers keep the files apart)
        out_h = gather_in_order_host(bufs, dst=0)
        assert (out_h == out) if rank == 0 else (out_h is None)
    if rank == 0:
        q.put([bytes(x) for x in out])
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_ordered_gather_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out == [bytes([p]) * (p * 37 % 11 + (0 if p == 3 else 1)) for p in range(7)]
