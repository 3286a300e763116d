#!/usr/bin/env python
"""Wall-clock of the drop-in ``Cropper.process_dir`` on a directory of JPEGs (decode / encode stay on the
host, cv2.imread / cv2.imwrite like the reference, cropper.py:554-609, utils.py:228-271).  With the compute path at
~750 img/s the directory run is bound by JPEG decode; the reference's own concurrency knob, ``num_processes`` (a
ThreadPool over batches, cropper.py:900-909), overlaps the host decode / encode of some batches with the GPU call of
another (cv2 releases the GIL; the GPU calls serialise on the one context).

    python profiles/process_dir_bench.py [--images 512] [--size 1024] [--batch 64]
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import sys
import tempfile
import time
from multiprocessing.pool import ThreadPool
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import cv2  # noqa: E402
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=512)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    from face_crop_plus_b200 import Cropper, synth
    base = synth.make_images(16, args.size, args.size, seed=1234)
    root = Path(tempfile.mkdtemp(prefix="fcp_dir_"))
    src = root / "in"
    src.mkdir()
    for i in range(args.images):
        cv2.imwrite(str(src / f"img{i:05d}.jpg"), cv2.cvtColor(base[i % 16], cv2.COLOR_RGB2BGR), [cv2.IMWRITE_JPEG_QUALITY, 90])
    files = sorted(os.listdir(src))
    # host I/O ceilings
    for threads in (1, os.cpu_count() or 1):
        t0 = time.perf_counter()
        with ThreadPool(threads) as pool:
            pool.map(lambda f: cv2.imread(str(src / f)), files)
        print(json.dumps({"what": "cv2.imread only", "threads": threads, "images_per_sec": args.images / (time.perf_counter() - t0)}), flush=True)
    sds = {"det": synth.make_state_dict("retinaface", 0, class_bias=4.8), "par": synth.make_state_dict("bisenet", 0)}
    for procs in (1, 4, 16):
        cr = Cropper(output_size=256, output_format="jpg", resize_size=args.size, strategy="largest", det_threshold=0.6,
                     mask_groups={"skin": [1]}, batch_size=args.batch, num_processes=procs, device="cuda:0", state_dicts=sds)
        out = root / f"out{procs}"
        cr.process_dir(str(src), str(out), desc=None)          # warm-up (weights, arena, page cache)
        shutil.rmtree(out)
        t0 = time.perf_counter()
        cr.process_dir(str(src), str(out), desc=None)
        dt = time.perf_counter() - t0
        n_out = sum(1 for _ in out.rglob("*.jpg"))
        print(json.dumps({"what": "Cropper.process_dir (detect+align+parse+save)", "num_processes": procs, "batch_size": args.batch,
                          "images": args.images, "files_written": n_out, "seconds": dt, "images_per_sec": args.images / dt}), flush=True)
    shutil.rmtree(root)


if __name__ == "__main__":
    main()
