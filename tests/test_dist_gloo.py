"""CPU, world_size 2 over gloo: the multi-rank plumbing used by bench.py (sharding + timing reduction)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import isg_b200  # noqa: F401
        from isg_b200 import dist as idist
        lo, hi = idist.shard_range(64, rank, world)
        t = idist.reduce_max(1.0 + rank)             # slowest rank defines the step time
        s = idist.reduce_sum(hi - lo)
        g = idist.gather_ints([rank, hi - lo])
        # sharded decode of a global batch of 5 "images": every rank decodes its shard, rank 0 gets all results in order
        kp = torch.arange(5, dtype=torch.float32).view(5, 1, 1, 1)
        outs = ((kp, kp.repeat(1, 4, 1, 1), None), torch.zeros(5, 3, 4), torch.zeros(5, 3, 2), torch.zeros(1, 3, 4))
        seen = []

        def fake_decode(inputs, o, infos, transforms, cfg, device):
            seen.append((inputs.shape[0], o[0][0].shape[0], o[1].shape[0], o[2].shape[0], o[3].shape[0], len(infos)))
            return [[("img", int(v), info)] * int(v) for v, info in zip(o[0][0].flatten().tolist(), infos)]   # ragged
        full = idist.decode_output_sharded(torch.zeros(5, 3, 1, 1), outs, ["i%d" % i for i in range(5)], None, None, "cpu",
                                           decode_fn=fake_decode)
        q.put((rank, lo, hi, t, s, g, full, seen))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 32), (32, 64)]
    assert all(r[3] == 2.0 and r[4] == 64.0 for r in res)
    assert res[0][5] == [[0, 32], [1, 32]]
    assert res[1][6] is None and res[0][6] == [[("img", v, "i%d" % v)] * v for v in range(5)]
    assert res[0][7] == [(3, 3, 3, 3, 1, 3)] and res[1][7] == [(2, 2, 2, 2, 1, 2)]


def test_shard_range_covers_everything():
    import isg_b200  # noqa: F401
    from isg_b200 import dist as idist
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [idist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        idist.shard_range(4, 2, 2)
    assert idist.reduce_max(3.5) == 3.5 and idist.gather_ints([1, 2]) == [[1, 2]]
    idist.barrier()                                                           # no process group: a no-op
