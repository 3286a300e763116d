"""CPU (gloo, world_size 2): batch sharding + the single metadata all-gather reproduce the single-process result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from face_crop_plus_b200 import distributed as D


def fake_local(images):
    """Deterministic stand-in for the GPU pipeline: one 'face' per image whose mean is even, derived from pixel data."""
    idx = [i for i, im in enumerate(images) if int(im.sum()) % 2 == 0]
    lms = np.stack([np.full((5, 2), float(images[i].mean()), np.float32) + np.arange(10, dtype=np.float32).reshape(5, 2)
                    for i in idx]) if idx else np.zeros((0, 5, 2), np.float32)
    mats = np.stack([np.arange(6, dtype=np.float64).reshape(2, 3) * (1 + images[i].astype(np.float64).sum()) for i in idx]) \
        if idx else np.zeros((0, 2, 3))
    return dict(landmarks=lms, indices=idx, matrices=mats, valid=np.ones(len(idx), bool))


def worker(rank, world, port, images, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = D.process_sharded(fake_local, images, rank, world, capacity_per_rank=len(images))
    q.put((rank, out["landmarks"], out["indices"], out["matrices"], out["valid"], out["owner"], out["shard"]))
    dist.barrier()
    dist.destroy_process_group()


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_ranges_cover_batch():
    for n in (0, 1, 7, 256):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_two_rank_gather_equals_single_process():
    rng = np.random.default_rng(0)
    images = rng.integers(0, 256, (7, 8, 8, 3), dtype=np.uint8)
    ref = fake_local(images)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, images, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, lms, idx, mats, valid, owner, shard in results:
        assert idx == ref["indices"]                                     # global image indices, ascending
        assert np.array_equal(lms, ref["landmarks"]) and np.array_equal(mats, ref["matrices"]) and valid.all()
        assert owner.tolist() == [0 if i < 4 else 1 for i in idx]        # rank 0 owns images 0..3, rank 1 owns 4..6
        assert shard == D.shard_range(7, rank, 2)


def test_single_process_path_and_capacity():
    images = np.zeros((3, 4, 4, 3), np.uint8)
    out = D.process_sharded(fake_local, images, 0, 1, capacity_per_rank=3)
    assert out["indices"] == [0, 1, 2]
    try:
        D.gather_records(torch.zeros((5, D.RECORD), dtype=torch.float64), capacity=2)
    except ValueError:
        pass
    else:
        raise AssertionError("capacity overflow must raise")
