"""CPU, gloo, world_size 2: the batch scatter / result gather indexing of the data-parallel path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, B, L, T, q):
    import sys
    sys.path.insert(0, ROOT)
    import viet_asr_b200  # noqa: F401
    from viet_asr_b200 import dist as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dev = torch.device("cpu")
        wave = length = None
        if rank == 0:
            wave = torch.arange(B * L, dtype=torch.float32).reshape(B, L)
            length = torch.arange(B, dtype=torch.int64) + 100
        w, ln = D.scatter_batch(wave, length, B, L, dev)
        s, e = D.shard_bounds(B, world)[rank]
        assert w.shape == (e - s, L) and ln.tolist() == [100 + i for i in range(s, e)]
        if e > s:
            assert w[0, 0].item() == s * L
        # int16 PCM shards (audio-ingest path): same indexing, half the bytes
        pcm = (torch.arange(B * L, dtype=torch.int32).reshape(B, L) % 30000).to(torch.int16) if rank == 0 else None
        wp, lp = D.scatter_batch(pcm, length, B, L, dev, dtype=torch.int16)
        assert wp.dtype == torch.int16 and wp.shape == (e - s, L) and lp.tolist() == ln.tolist()
        if e > s:
            assert wp[0, 1].item() == (s * L + 1) % 30000
        # "transcribe": ids = utterance index repeated, length = index
        ids = torch.stack([torch.full((T,), i, dtype=torch.int32) for i in range(s, e)]) if e > s else torch.empty((0, T), dtype=torch.int32)
        n = torch.arange(s, e, dtype=torch.int32)
        gi, gl = D.gather_results(ids, n, B)
        if rank == 0:
            assert gi.shape == (B, T) and gl.tolist() == list(range(B))
            assert gi[:, 0].tolist() == list(range(B))
        else:
            assert gi is None and gl is None
        q.put((rank, "ok"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [5, 8, 1])
def test_scatter_gather_world2(B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, 7, 3, q)) for r in range(2)]
    for p in procs: p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got == [(0, "ok"), (1, "ok")]


def test_shard_bounds_cover_and_balance():
    import viet_asr_b200  # noqa: F401
    from viet_asr_b200.dist import shard_bounds
    for n in (0, 1, 7, 256, 1024, 1025):
        for w in (1, 2, 4, 8):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1
