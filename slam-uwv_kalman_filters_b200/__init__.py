"""B200-native batched manifold-UKF engine (uwvkf): host-side Python mirror of the C ABI in include/uwvkf.h.

The product is the CUDA library `libuwvkf.so` (built in-tree by `__graft_entry__.build()` or
`make -C slam-uwv_kalman_filters_b200/csrc`).  This package only binds it (ctypes) and mirrors the reference's
filter-class interface; it contains no CPU implementation of the filters and raises if the library is missing.
"""
from .capi import (PoseSampleStream, STREAM_ROTATION_RATE, PoseUKFBatch, VelocityUKFBatch, BottomUKFBatch, IndirectPoseUKFBatch, UwvkfError, load_library, library_path, POSE_XY, POSE_Z, POSE_PRESSURE,  # noqa: F401
                   POSE_ACCELERATION, POSE_VELOCITY, POSE_WATER_VELOCITY, POSE_GEOGRAPHIC, POSE_DELAYED_XY, POSE_BODY_EFFORTS, POSE_BODY_EFFORTS_VELOCITY_ONLY,
                   ST_NOT_SPD, ST_NAN_MEASUREMENT, ST_GATE_REJECTED, ST_MEAN_NOT_CONVERGED, HOST, DEVICE)
