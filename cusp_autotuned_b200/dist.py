"""torch.distributed plumbing for the row-partitioned operator: rendezvous, and
hand-over of the NCCL unique id to the engine's own communicator
(b200sp_comm_init).  Data-path collectives (halo send/recv, scalar all-reduce)
are issued by libb200sp itself on the compute stream — not through torch."""
from __future__ import annotations

import os

import torch
import torch.distributed as td


def init_process_group_from_env(backend: str | None = None):
    """one process per GPU, launched by torch.distributed.run; returns (rank, world, local_rank)"""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not td.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            td.init_process_group(backend, rank=rank, world_size=world,
                                  device_id=torch.device("cuda", local))
        else:
            td.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0) -> bytes:
    """rank `src` supplies `payload`; everyone returns it (works on gloo and nccl)"""
    if not td.is_initialized() or td.get_world_size() == 1:
        return payload
    dev = torch.device("cuda", torch.cuda.current_device()) if td.get_backend() == "nccl" else torch.device("cpu")
    t = torch.zeros(nbytes, dtype=torch.uint8)
    if td.get_rank() == src:
        t = torch.frombuffer(bytearray(payload), dtype=torch.uint8).clone()
    t = t.to(dev)
    td.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def init_engine_comm(handle, rank: int, world: int):
    """create the engine's NCCL communicator (unique id from rank 0)"""
    if world <= 1:
        return
    uid = handle.comm_unique_id() if rank == 0 else None
    uid = broadcast_bytes(uid, 128, src=0)
    handle.comm_init(uid, world, rank)
