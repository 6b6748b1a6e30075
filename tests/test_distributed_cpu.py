"""world_size-2 gloo test (CPU) of the host logic of the multi-GPU path: file sharding, bucket ownership, id broadcast,
max-over-ranks timing, gathering of the per-rank file lists."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hyperspace_b200 import distributed as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    files = [f"part-{i:05d}.parquet" for i in range(7)]
    mine = D.shard_files(files, rank, world)
    uid = D.broadcast_unique_id(dist, lambda: bytes(range(128)), rank)
    assert uid == bytes(range(128))
    assert D.max_over_ranks(dist, float(rank + 1)) == float(world)
    owned = D.buckets_of_rank(rank, world, 200)
    written = [f"part-{b:05d}-u_{b:05d}.c000.parquet" for b in owned]
    everything = D.gather_file_lists(dist, written, world)
    all_shards = [None] * world
    dist.all_gather_object(all_shards, mine)
    if rank == 0:
        assert [f for s in all_shards for f in s] == files                      # shards partition the file list in order
        assert sorted(int(f.rsplit("_", 1)[1].split(".")[0]) for f in everything) == list(range(200))  # every bucket once
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    dist.destroy_process_group()


def test_two_rank_host_logic(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "ok").exists()


def test_bucket_ownership_matches_exchange_kernel():
    for world in (1, 2, 4, 8):
        seen = sorted(b for r in range(world) for b in D.buckets_of_rank(r, world, 200))
        assert seen == list(range(200))
        assert all(D.owner_of_bucket(b, world) == b % world for b in range(200))
