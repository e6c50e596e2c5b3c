"""Host-side logic of the multi-GPU path on CPU: two gloo ranks shard a stream round-robin,
all-gather their key records and agree on the newest writer per key."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from edge_fuse_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_local, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pos = sharding.shard_positions(rank, world, n_local, base=100)
    # addresses: page = position % 24 -> the same key is written by both ranks at different times
    u = np.full(n_local, 77, dtype=np.uint64)
    l = pos % np.uint64(24)
    lens = np.where(pos % np.uint64(11) == 0, -1, 1000 + pos.astype(np.int64))   # some chunks store nothing
    rec = sharding.pack_records(u, l, pos, rank, lens)
    gathered = sharding.all_gather_records(torch.from_numpy(rec))
    rows = sharding.remote_rows(gathered, rank)
    q.put((rank, gathered.numpy().copy(), rows.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange():
    world, n_local = 2, 40
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_local, q)) for r in range(world)]
    [p.start() for p in procs]
    got = {}
    for _ in range(world):
        rank, gathered, rows = q.get(timeout=120)
        got[rank] = (gathered, rows)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    g0, g1 = got[0][0], got[1][0]
    assert (g0 == g1).all() and g0.shape == (world * n_local, 4)          # everyone sees the same gather
    u, l, seq, owner, length = sharding.unpack_records(g0)
    assert set(np.uint64(seq).tolist()) == set(range(100, 100 + world * n_local))   # round-robin covers the stream
    assert ((np.uint64(seq) - 100) % world == owner).all()
    assert (length[np.uint64(seq) % 11 == 0] == -1).all() and (length[np.uint64(seq) % 11 != 0] >= 0).all()
    for rank in range(world):
        rows = got[rank][1]
        _, _, _, ro, rl = sharding.unpack_records(rows)
        assert (ro != rank).all() and (rl >= 0).all()
        assert len(rows) == int(((owner != rank) & (length >= 0)).sum())
    # newest writer per key == what a sequential pass over the global stream leaves behind
    best = sharding.resolve_newest(g0)
    seqs = np.uint64(seq)
    order = np.argsort(seqs)
    sequential = {}
    for i in order:
        if length[i] >= 0:
            sequential[(77, int(np.uint64(l[i])))] = (int(seqs[i]), int(owner[i]))
    assert best == sequential and len(best) == 24


def test_pack_unpack_roundtrip():
    u = np.array([2**63 + 5, 1], dtype=np.uint64)
    l = np.array([7, 2**44 - 1], dtype=np.uint64)
    rec = sharding.pack_records(u, l, np.array([9, 2**40], dtype=np.uint64), 3, np.array([-1, 65794]),
                                rec_off=np.array([4096, 179 * 2**30 + 48], dtype=np.uint64))
    uu, ll, ss, oo, nn = sharding.unpack_records(rec)
    assert (np.uint64(uu) == u).all() and (np.uint64(ll) == l).all()
    assert oo.tolist() == [3, 3] and nn.tolist() == [-1, 65794] and np.uint64(ss).tolist() == [9, 2**40]
    # the location travels with rows that stored something (16-byte units, 34 bits: arenas up to 256 GiB)
    assert sharding.unpack_locations(rec).tolist() == [0, 179 * 2**30 + 48]
    rec7 = sharding.pack_records(u, l, np.array([1, 2], dtype=np.uint64), 7, np.array([0, 2**21]))
    assert sharding.unpack_records(rec7)[3].tolist() == [7, 7] and sharding.unpack_records(rec7)[4].tolist() == [0, 2**21]
    assert (sharding.shard_positions(1, 4, 3, 10) == np.array([11, 15, 19], dtype=np.uint64)).all()
