"""Data-parallel sharding of utterances over the GPUs of one box.

Utterances are independent (BatchNorm is in eval mode, feature normalisation is
per utterance), so the path shards with no data-path collective: every rank runs
the whole model on a contiguous slice of the batch.  The only exchanges are the
batch scatter (rank 0 -> all: waveforms + lengths) and the result gather
(all -> rank 0: collapsed ids + lengths), done with torch.distributed
(NCCL over NVLink on GPUs; gloo in the CPU tests).  This replaces the
reference's per-tensor all_gather of padded eval results
(nemo/backends/pytorch/actions.py:444-478, 594-612, 784-802).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced split of n utterances: the first n % world ranks get one extra."""
    base, extra = divmod(n, world)
    out, start = [], 0
    for r in range(world):
        cnt = base + (1 if r < extra else 0)
        out.append((start, start + cnt))
        start += cnt
    return out


def scatter_batch(wave: Optional[torch.Tensor], length: Optional[torch.Tensor], B: int, L: int,
                  device: torch.device, src: int = 0, dtype: torch.dtype = torch.float32):
    """Rank `src` holds wave [B, L] (`dtype`: float32, or int16 PCM for the audio-ingest path - half the bytes on the
    wire) and length [B] i64; every rank returns its slice.
    Uneven shards are padded to the largest shard for the collective and trimmed after."""
    world, rank = dist.get_world_size(), dist.get_rank()
    bounds = shard_bounds(B, world)
    per = max(e - s for s, e in bounds)
    my_n = bounds[rank][1] - bounds[rank][0]
    # neither NCCL nor gloo moves int16: PCM travels as its bytes (uint8 view), any other dtype as it is
    as_bytes = dtype == torch.int16
    w_out = torch.empty((per, L), dtype=dtype, device=device)
    w_wire = w_out.view(torch.uint8) if as_bytes else w_out
    l_out = torch.empty((per,), dtype=torch.int64, device=device)
    if rank == src:
        if wave.dtype != dtype:
            raise ValueError(f"scatter_batch: wave is {wave.dtype}, dtype argument says {dtype}")
        w_list, l_list = [], []
        for s, e in bounds:
            w = torch.zeros((per, L), dtype=dtype, device=device)
            ln = torch.full((per,), L, dtype=torch.int64, device=device)
            w[: e - s] = wave[s:e].to(device)
            ln[: e - s] = length[s:e].to(device)
            w_list.append(w.view(torch.uint8) if as_bytes else w); l_list.append(ln)
        dist.scatter(w_wire, w_list, src=src)
        dist.scatter(l_out, l_list, src=src)
    else:
        dist.scatter(w_wire, None, src=src)
        dist.scatter(l_out, None, src=src)
    return w_out[:my_n], l_out[:my_n]


def gather_results(out_ids: torch.Tensor, out_len: torch.Tensor, B: int, dst: int = 0):
    """Every rank holds out_ids [n_r, T] i32 and out_len [n_r] i32; rank `dst` returns the
    batch-ordered [B, T] / [B] tensors, other ranks (None, None)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    bounds = shard_bounds(B, world)
    per = max(e - s for s, e in bounds)
    T = out_ids.shape[1]
    ids_p = torch.full((per, T), -1, dtype=torch.int32, device=out_ids.device)
    len_p = torch.zeros((per,), dtype=torch.int32, device=out_ids.device)
    n = out_ids.shape[0]
    ids_p[:n] = out_ids
    len_p[:n] = out_len
    if rank == dst:
        ids_l = [torch.empty_like(ids_p) for _ in range(world)]
        len_l = [torch.empty_like(len_p) for _ in range(world)]
        dist.gather(ids_p, ids_l, dst=dst)
        dist.gather(len_p, len_l, dst=dst)
        ids = torch.cat([t[: e - s] for t, (s, e) in zip(ids_l, bounds)], dim=0)
        lens = torch.cat([t[: e - s] for t, (s, e) in zip(len_l, bounds)], dim=0)
        return ids, lens
    dist.gather(ids_p, None, dst=dst)
    dist.gather(len_p, None, dst=dst)
    return None, None
