"""Host-side helpers for the one-process-per-GPU launch (torchrun): how the source files are split across ranks, which
rank owns which bucket, and the rendezvous of the NCCL unique id.  torch.distributed is plumbing only: the data path's
single exchange is the grouped ncclSend/ncclRecv all-to-all inside hs_create_index (csrc/exchange.cu), the analogue of
the shuffle behind ``indexData.repartition(numBuckets, indexedColumns)``
(src/main/scala/com/microsoft/hyperspace/index/covering/CoveringIndex.scala:60)."""
from __future__ import annotations

import os
import subprocess
from typing import List, Optional, Sequence, TypeVar

T = TypeVar("T")


def shard_files(files: Sequence[T], rank: int, world: int) -> List[T]:
    """Contiguous, near-equal split of the source file list (the map side's input splits)."""
    n = len(files)
    return list(files[rank * n // world:(rank + 1) * n // world])


def owner_of_bucket(bucket: int, world: int) -> int:
    """Must match exchange.cu: owner(b) = b mod world."""
    return bucket % world


def buckets_of_rank(rank: int, world: int, num_buckets: int) -> List[int]:
    return [b for b in range(num_buckets) if owner_of_bucket(b, world) == rank]


def broadcast_unique_id(dist, make_id, rank: int) -> bytes:
    """Rank 0 creates the 128-byte NCCL id (hs_comm_unique_id); everyone receives it."""
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def max_over_ranks(dist, value: float, device=None) -> float:
    import torch

    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_file_lists(dist, local_files: List[str], world: int) -> List[str]:
    """Every rank wrote the files of the buckets it owns; rank-major concatenation = the index content."""
    box = [None] * world
    dist.all_gather_object(box, local_files)
    return [f for part in box for f in part]


def _parse_cpulist(text: str) -> List[int]:
    out: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            out.extend(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    return out


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pins this process to the CPUs of the NUMA node the GPU hangs off, BEFORE any pinned host memory is allocated, so that
    the file images a rank stages (hs_stage_sources) and drains (hs_pending_wait) live in memory local to its GPU's PCIe
    root.  With eight ranks copying in both directions at once, images on the far socket make every byte cross the
    inter-socket link as well.  What ``numactl --cpunodebind --membind`` does for a Spark executor; a no-op (returns None)
    when the topology cannot be read.  Returns the node."""
    try:
        bus = subprocess.check_output(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(device_index)],
                                      text=True, timeout=20).strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set(_parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())) & os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None
