"""Batch-shard data parallelism for the hot path: one process per GPU, images split contiguously across ranks, and
exactly ONE collective per batch — an all-gather of fixed-size per-face metadata records.

The reference has no multi-device code (one ``torch.device`` per ``Cropper``, cropper.py:336-337); images are
independent end to end, so nothing else ever crosses devices: crops / labels / masks stay on the owning GPU.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

RECORD = 20   # float64 per face: landmarks[10], global image index, matrix[6], valid, score slot, reserved


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous shard [lo, hi) of ``n`` images owned by ``rank`` (sizes differ by at most one)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_records(landmarks, indices, matrices, valid, image_offset: int) -> torch.Tensor:
    """Local per-face results -> float64 [F, RECORD] records with GLOBAL image indices."""
    f = len(indices)
    rec = torch.zeros((f, RECORD), dtype=torch.float64)
    if f:
        rec[:, :10] = torch.as_tensor(np.asarray(landmarks, dtype=np.float64).reshape(f, 10))
        rec[:, 10] = torch.as_tensor(np.asarray(indices, dtype=np.float64)) + image_offset
        rec[:, 11:17] = torch.as_tensor(np.asarray(matrices, dtype=np.float64).reshape(f, 6))
        rec[:, 17] = torch.as_tensor(np.asarray(valid, dtype=np.float64))
    return rec


def gather_records(rec: torch.Tensor, capacity: int, group=None, device=None) -> dict:
    """The one collective: all-gather ``capacity`` records (+ the count) from every rank.

    Returns the batch-global ``landmarks`` f32 [F,5,2], ``indices`` list[int] (ascending, like RetinaFace.predict),
    ``matrices`` f64 [F,2,3], ``valid`` bool [F] and ``owner`` (rank holding each face's crop) — identical on all ranks.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    f = rec.shape[0]
    if f > capacity:
        raise ValueError(f"{f} faces exceed the all-gather capacity {capacity}")
    buf = torch.zeros((capacity + 1, RECORD), dtype=torch.float64, device=device)
    buf[:f] = rec.to(buf.device)
    buf[capacity, 0] = f
    if world > 1:
        out = torch.empty((world * (capacity + 1), RECORD), dtype=torch.float64, device=buf.device)
        dist.all_gather_into_tensor(out, buf, group=group)
    else:
        out = buf
    return unpack_records(out.cpu().view(world, capacity + 1, RECORD), capacity)


def unpack_records(out, capacity: int) -> dict:
    """[world, capacity + 1, RECORD] float64 blocks (torch tensor or numpy array; slot ``capacity`` = the rank's face count)
    -> the batch-global metadata dict of :func:`gather_records`."""
    out = torch.as_tensor(out).cpu()
    world = out.shape[0]
    parts, owner = [], []
    for r in range(world):
        k = int(out[r, capacity, 0].item())
        parts.append(out[r, :k])
        owner += [r] * k
    allrec = torch.cat(parts) if parts else torch.zeros((0, RECORD), dtype=torch.float64)
    return dict(landmarks=allrec[:, :10].to(torch.float32).view(-1, 5, 2).numpy(), indices=allrec[:, 10].long().tolist(),
                matrices=allrec[:, 11:17].view(-1, 2, 3).numpy(), valid=allrec[:, 17].bool().numpy(),
                owner=np.array(owner, dtype=np.int64))


def init_comm(ctx, group=None) -> None:
    """Creates the CUDA library's own NCCL communicator for ``ctx`` (one context per rank / GPU): rank 0 makes the NCCL
    unique id, ``torch.distributed`` (any backend) carries its 128 bytes to the other ranks, then every rank joins.
    After this, ``ctx.set_gather(buffer, cap, base)`` makes every ``ctx.pipeline`` call all-gather its face records on the
    device, overlapped with the parser — no host round trip (``fcp_set_gather``)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ids = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0, group=group)
    ctx.comm_init(rank, world, ids[0])


def process_sharded(run_local, images, rank: int, world: int, capacity_per_rank: int, group=None, device=None) -> dict:
    """Shards ``images`` (u8 [N,H,W,3]) by batch, runs ``run_local(shard) -> dict(landmarks, indices, matrices, valid, ...)``
    on this rank's shard and all-gathers the metadata.  ``local`` in the result holds this rank's full local outputs."""
    lo, hi = shard_range(len(images), rank, world)
    local = run_local(images[lo:hi])
    rec = pack_records(local["landmarks"], local["indices"], local["matrices"], local["valid"], lo)
    out = gather_records(rec, capacity_per_rank, group, device)
    out["local"] = local
    out["shard"] = (lo, hi)
    return out
