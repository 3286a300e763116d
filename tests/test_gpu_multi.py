"""Two GPUs, two processes: the batch-sharded path with the library's own NCCL all-gather (fcp_comm_init / fcp_set_gather)
reproduces the single-GPU result.  Skipped on boxes with fewer than two GPUs (run it with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from face_crop_plus_b200 import _abi, distributed as D, synth
    from face_crop_plus_b200.landmarks import landmarks_target
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)          # carries only the 128-byte NCCL id
    torch.cuda.set_device(rank)
    ctx = _abi.Context(rank)
    ctx.load_state_dict(_abi.MODEL_RETINAFACE, synth.make_state_dict("retinaface", 0, class_bias=4.0))
    ctx.load_state_dict(_abi.MODEL_BISENET, synth.make_state_dict("bisenet", 0))
    D.init_comm(ctx)
    imgs = synth.make_images(7, 256, 320, seed=2000)
    tgt = landmarks_target((256, 256), 0.65)
    lo, hi = D.shard_range(len(imgs), rank, world)
    cap = 8
    buf = torch.zeros((world, cap + 1, 20), dtype=torch.float64, device=f"cuda:{rank}")
    ctx.set_gather(buf, cap, lo)
    local = ctx.pipeline(np.ascontiguousarray(imgs[lo:hi]), None, tgt, (256, 256), 0.6, 0.4, "largest")
    ctx.set_gather(None)
    got = D.unpack_records(buf, cap)
    if rank == 0:
        whole = ctx.pipeline(imgs, None, tgt, (256, 256), 0.6, 0.4, "largest")
        q.put((got["indices"], got["landmarks"], got["matrices"], got["owner"].tolist(), whole["indices"].tolist(), whole["landmarks"],
               whole["matrices"], local["count"]))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_sharded_pipeline_with_nccl_gather_equals_single_gpu():
    import torch.multiprocessing as mp
    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    port = _free_port()
    procs = [mpctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    idx, lms, mats, owner, ref_idx, ref_lms, ref_mats, n0 = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert idx == ref_idx                                    # global image indices, rank order == image order
    assert np.array_equal(lms, ref_lms) and np.array_equal(mats, ref_mats)      # the path is position independent: bit-identical
    assert owner == [0] * n0 + [1] * (len(idx) - n0)
