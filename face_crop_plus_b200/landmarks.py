"""Target-landmark template of the align step.

Mirrors ``utils.py:13-19`` (the 5-point normalised template, domain data) and
``cropper.py:392-439`` (scaling by output size / face factor in float32).
"""
from __future__ import annotations

import numpy as np

#: normalised (x, y) of left eye, right eye, nose tip, left / right mouth corner (observer's view)
STANDARD_LANDMARKS_5 = np.float32([
    [0.31556875000000000, 0.4615741071428571],
    [0.68262291666666670, 0.4615741071428571],
    [0.50026249999999990, 0.6405053571428571],
    [0.34947187500000004, 0.8246919642857142],
    [0.65343645833333330, 0.8246919642857142],
])


def landmarks_target(output_size: tuple[int, int], face_factor: float, num_std_landmarks: int = 5) -> np.ndarray:
    """float32 [5,2] target landmarks; same operation order as ``cropper.py:431-436``."""
    if num_std_landmarks != 5:
        raise ValueError(f"Unsupported number of standard landmarks for estimating alignment transform matrix: "
                         f"{num_std_landmarks}.")
    std = STANDARD_LANDMARKS_5.copy()
    std[:, 0] *= output_size[0] * face_factor
    std[:, 1] *= output_size[1] * face_factor
    std[:, 0] += (1 - face_factor) * output_size[0] / 2
    std[:, 1] += (1 - face_factor) * output_size[1] / 2
    return std
