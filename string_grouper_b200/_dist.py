"""Multi-GPU sharding of the hot path: one process per GPU (torchrun), torch.distributed for plumbing.

The block loop of StringGrouper._build_matches (/root/reference/string_grouper/string_grouper.py:734-750)
is data-parallel over LEFT row blocks (their results are only `vstack`ed, :750); right blocks need a
per-row merge (:746), so the right side is never split across GPUs.  Rank g owns the left rows
`define_chunks(n_left, world)[g]` (same ceil-sized chunking as :721-722), runs K2 on them against the
full right matrix, and the small per-rank top-n lists are all-gathered (variable length) so that every
rank holds the complete match list, in rank order == row order.  K1 runs redundantly on every rank
(the packed corpus is a few tens of MB); the only collective on the data path is the result gather.
"""
import numpy as np


def _dist():
    import torch.distributed as dist
    return dist


def world():
    """(rank, world_size) of the default process group, (0, 1) when not running distributed."""
    try:
        dist = _dist()
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def shard_range(n_rows, rank, world_size):
    """Rows of chunk `rank` when `n_rows` are split into `world_size` ceil-sized consecutive chunks
    (define_chunks, string_grouper.py:714-722); trailing chunks may be empty."""
    chunk = -(-int(n_rows) // int(world_size)) if n_rows > 0 else 0
    lo = min(rank * chunk, n_rows)
    hi = min(lo + chunk, n_rows)
    return lo, hi


def allgather_varlen(tensors, group=None):
    """All-gather a tuple of equally long 1-D tensors whose length differs per rank.

    Returns the tuple of concatenations in rank order.  Works on CPU tensors (gloo) and CUDA tensors
    (nccl over NVLink); lengths are exchanged first, payloads are padded to the longest.
    """
    import torch
    dist = _dist()
    ws = dist.get_world_size(group)
    n = int(tensors[0].numel())
    dev = tensors[0].device
    lens = torch.zeros(ws, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(lens, torch.tensor([n], dtype=torch.int64, device=dev), group=group)
    lens = lens.cpu().tolist()
    cap = max(max(lens), 1)
    out = []
    for x in tensors:
        pad = torch.zeros(cap, dtype=x.dtype, device=dev)
        pad[:n] = x[:n]
        buf = torch.empty(ws * cap, dtype=x.dtype, device=dev)
        dist.all_gather_into_tensor(buf, pad, group=group)
        out.append(torch.cat([buf[r * cap:r * cap + lens[r]] for r in range(ws)]))
    return tuple(out), lens


def gather_matches(shape, row, col, score, nnz, max_row):
    """Combine the per-rank top-n lists of DeviceMatches into the global list on every rank."""
    import torch
    dist = _dist()
    (g_row, g_col, g_score), lens = allgather_varlen((row[:nnz], col[:nnz], score[:nnz]))
    mx = torch.tensor([int(max_row)], dtype=torch.int64, device=row.device)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    return g_row, g_col, g_score, int(sum(lens)), int(mx.item())


def shard_vectorise(n_bytes_total):
    """Whether K1 should be sharded over the ranks (two-Series inputs only).  SG_B200_SHARD_VECTORISE = 0 / 1 /
    auto (default): auto shards once the packed corpus exceeds 256 MB — below that every rank vectorising
    everything costs a few milliseconds and needs no collective at all."""
    import os
    mode = os.environ.get("SG_B200_SHARD_VECTORISE", "auto").lower()
    if mode in ("1", "true", "yes", "on"):
        return True
    if mode in ("0", "false", "no", "off"):
        return False
    return n_bytes_total > (256 << 20)


def allreduce_sum_(tensor):
    """In-place sum over the ranks (the int32 document-frequency table of K1: ncclAllReduce over NVLink)."""
    dist = _dist()
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


def allgather_csr_rows(row_len, indices, vals):
    """All-gather a row-sharded CSR (the right matrix after a sharded K1).

    row_len: int64 lengths of the local rows; indices: int32 [local nnz]; vals: tuple of value tensors
    [local nnz].  Returns (indptr int64 [n_total+1], indices, vals) of the concatenation in rank order.
    """
    import torch
    (all_len,), _ = allgather_varlen((row_len,))
    gathered, _ = allgather_varlen((indices,) + tuple(vals))
    indptr = torch.zeros(all_len.numel() + 1, dtype=torch.int64, device=row_len.device)
    torch.cumsum(all_len, 0, out=indptr[1:])
    return indptr, gathered[0], gathered[1:]


def shard_offsets(n_rows, world_size):
    return np.array([shard_range(n_rows, r, world_size)[0] for r in range(world_size)] + [n_rows], dtype=np.int64)
