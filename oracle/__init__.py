"""CPU oracle for the face-crop-plus hot path — TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the shipped path
(``face_crop_plus_b200``) never does and fails loudly without its CUDA library.

What it is: a restatement, in torch-CPU functional ops / numpy, of the reference
algorithm for detect -> align -> (enhance) -> parse, each function citing the
reference file:line it follows.  The convolution arithmetic itself lives in
third-party libraries the reference merely calls (torch/ATen, torchvision
resnet50, OpenCV ``estimateAffine*2D`` / ``warpAffine`` — all unpinned in the
reference's ``setup.py:36-42``; versions used here: torch 2.11.0, torchvision
0.26.0, opencv 4.13.0).

Parity pin: the reference ships no tests or golden vectors, so
the oracle is pinned against *outputs of the unmodified reference run in the
build container* on seeded synthetic weights/images: ``oracle/make_golden.py``
imports ``/root/reference/src`` and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every oracle function against them.
"""
