"""Multi-GPU plumbing for the cachemap put path (option B).

Chunk k of the global stream belongs to rank k mod world (round-robin); every rank encodes and
stores its own chunks with no data-path collective.  What is replicated is the key index: after
each batch the ranks all-gather one fixed-size record per stored chunk — {address u, address l,
global stream position, owner rank | stored length} = 32 bytes — over NCCL (NVLink / NVSwitch) and
import the other ranks' records into their table replica (cmb200_import_remote), where the highest
stream position per key wins, i.e. the outcome of the sequential reference.  torch.distributed is
only the transport; the table logic is in the CUDA library.
"""
from __future__ import annotations

import numpy as np

REC_WORDS = 4   # int64 words per record


def shard_positions(rank: int, world: int, n_local: int, base: int = 0) -> np.ndarray:
    """Global stream positions of this rank's chunks: base + rank, base + rank + world, ..."""
    return (np.uint64(base) + np.uint64(rank) + np.uint64(world) * np.arange(n_local, dtype=np.uint64))


def pack_records(u, l, seq, rank: int, lens) -> np.ndarray:
    """[n, 4] int64: u, l, seq, (rank << 32 | len & 0xffffffff).  len < 0 marks a chunk that stored
    nothing (rejected address, superseded inside its batch, dropped) and is ignored by importers."""
    n = len(u)
    rec = np.empty((n, REC_WORDS), dtype=np.int64)
    rec[:, 0] = np.asarray(u, dtype=np.uint64).view(np.int64)
    rec[:, 1] = np.asarray(l, dtype=np.uint64).view(np.int64)
    rec[:, 2] = np.asarray(seq, dtype=np.uint64).view(np.int64)
    rec[:, 3] = (np.int64(rank) << np.int64(32)) | (np.asarray(lens, dtype=np.int64) & np.int64(0xFFFFFFFF))
    return rec


def unpack_records(rec):
    """-> u, l, seq, owner, length (arrays; works on numpy arrays and torch tensors alike)."""
    u, l, seq, tail = rec[..., 0], rec[..., 1], rec[..., 2], rec[..., 3]
    owner = tail >> 32
    length = ((tail & 0xFFFFFFFF) ^ 0x80000000) - 0x80000000      # sign-extend the low 32 bits
    return u, l, seq, owner, length


def all_gather_records(rec_tensor, group=None):
    """One all-gather of this rank's [n, 4] int64 records -> [world * n, 4] (same device)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty((world * rec_tensor.shape[0], REC_WORDS), dtype=torch.int64, device=rec_tensor.device)
    try:
        dist.all_gather_into_tensor(out, rec_tensor.contiguous(), group=group)
    except (RuntimeError, NotImplementedError):                    # backends without the flat form
        parts = [torch.empty_like(rec_tensor) for _ in range(world)]
        dist.all_gather(parts, rec_tensor.contiguous(), group=group)
        out = torch.cat(parts, dim=0)
    return out


def remote_rows(gathered, rank: int):
    """Rows written by other ranks that actually stored something (torch tensor or numpy array)."""
    _, _, _, owner, length = unpack_records(gathered)
    keep = (owner != rank) & (length >= 0)
    return gathered[keep]


def resolve_newest(records: np.ndarray) -> dict:
    """Reference resolution on the host (tests): key (u, l) -> (seq, owner) of the highest seq."""
    best = {}
    u, l, seq, owner, length = unpack_records(records)
    for i in range(len(records)):
        if length[i] < 0:
            continue
        k = (int(np.uint64(u[i])), int(np.uint64(l[i])))
        s = int(np.uint64(seq[i]))
        if k not in best or s > best[k][0]:
            best[k] = (s, int(owner[i]))
    return best


def import_gathered(engine, gathered, rank: int) -> int:
    """Imports the other ranks' rows of an all-gathered record tensor (CUDA tensor) into `engine`'s
    index replica.  Returns the number of rows imported."""
    import torch
    rows = remote_rows(gathered, rank)
    n = int(rows.shape[0])
    if n == 0:
        return 0
    addr = rows[:, :2].contiguous()
    seq = rows[:, 2].contiguous()
    owner = (rows[:, 3] >> 32).to(torch.int32).contiguous()
    if rows.is_cuda:
        torch.cuda.current_stream(rows.device).synchronize()
        from .binding import lib, _check
        _check(lib().cmb200_import_remote(engine.h, n, addr.data_ptr(), owner.data_ptr(), seq.data_ptr(), 1),
               "cmb200_import_remote")
    else:
        a = addr.numpy().view(np.uint64)
        engine.import_remote(a[:, 0], a[:, 1], owner.numpy().view(np.uint32), seq.numpy().view(np.uint64))
    return n
