"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for the bootstrap only.
The data path (k x k all-reduces, SpMM halo exchange) runs on the library's own NCCL
communicator created from a unique id that rank 0 broadcasts here."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import partition
from .api import _check, lib, set_csr, set_csr_b, set_lr


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_comm(dist=None):
    """Create the library's NCCL communicator over the ranks of torch.distributed's default
    group (already initialised by the caller).  No-op for a single process."""
    rank, world, _ = env_rank()
    if world <= 1 or dist is None:
        return rank, world
    uid = np.zeros(128, dtype=np.uint8)
    if rank == 0:
        _check(lib().diaglib_b200_comm_unique_id(C.c_void_p(uid.ctypes.data)), "comm_unique_id")
    box = [uid.tobytes()]
    dist.broadcast_object_list(box, src=0)
    uid = np.frombuffer(box[0], dtype=np.uint8).copy()
    _check(lib().diaglib_b200_comm_init(rank, world, C.c_void_p(uid.ctypes.data)), "comm_init")
    return rank, world


def install_partitioned(gen_rows, n: int, rank: int, world: int, dist=None, metric_rows=None, lr_rows=None):
    """gen_rows(r0, r1) -> (rowptr, col_global, val, diag) for the owned rows.  Localises the
    columns, exchanges the needed ranges and installs matrix + halo plan.  Returns (r0, r1).
    metric_rows(r0, r1) -> (rowptr, col_global, val) optionally installs the metric B of the
    generalized problem.
    lr_rows(r0, r1) -> dict(apb, amb, spd, smd = (rowptr, col_global, val), aa_diag, sigma_diag)
    optionally installs the linear-response matrices.
    All installed matrices share ONE halo: its ranges are the union of what each of them
    references on the other ranks."""
    r0, r1 = partition.row_range(n, rank, world)
    rowptr, col, val, diag = gen_rows(r0, r1)
    if world == 1:
        set_csr(rowptr, col, val, diag)
        if metric_rows is not None:
            set_csr_b(*metric_rows(r0, r1))
        if lr_rows is not None:
            lr = lr_rows(r0, r1)
            set_lr(lr["apb"], lr["amb"], lr["spd"], lr["smd"], lr["aa_diag"], lr["sigma_diag"])
        return r0, r1
    metric = metric_rows(r0, r1) if metric_rows is not None else None
    lr = lr_rows(r0, r1) if lr_rows is not None else None
    shared = [partition.needed_ranges(col, n, rank, world)]
    if metric is not None:
        shared.append(partition.needed_ranges(metric[1], n, rank, world))
    if lr is not None:
        shared += [partition.needed_ranges(lr[k][1], n, rank, world) for k in ("apb", "amb", "spd", "smd")]
    needed = partition.union_ranges(*shared)
    all_needed = [None] * world
    dist.all_gather_object(all_needed, needed)
    col_loc, n_halo, recv = partition.localize(col, n, rank, world, needed)
    plan = partition.halo_plan(recv, all_needed, n, rank, world)
    set_csr(rowptr, col_loc, val, diag, n_halo=n_halo, halo_plan=plan)
    if metric is not None:
        b_loc, _, _ = partition.localize(metric[1], n, rank, world, needed)
        set_csr_b(metric[0], b_loc, metric[2], n_halo=n_halo)
    if lr is not None:
        loc = {k: (lr[k][0], partition.localize(lr[k][1], n, rank, world, needed)[0], lr[k][2])
               for k in ("apb", "amb", "spd", "smd")}
        set_lr(loc["apb"], loc["amb"], loc["spd"], loc["smd"], lr["aa_diag"], lr["sigma_diag"], n_halo=n_halo)
    return r0, r1
