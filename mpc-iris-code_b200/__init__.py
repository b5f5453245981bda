"""B200-native matching backend for recmo/mpc-iris-code (distances + denominators hot path).

Host-side mirror of the reference's engine API (src/lib.rs:28-94) over the C ABI in
include/iris_b200.h.  The directory is named after the reference (`mpc-iris-code_b200/`);
import it as `mpc_iris_code_b200` (a shim package at the repo root extends its path here).

There is no CPU fallback: every compute call goes through libiris_b200.so and fails loudly
(IrisError / OSError) if the library or a CUDA device is missing.
"""
from .api import (  # noqa: F401
    BITS,
    COLS,
    Cluster,
    cluster_partition,
    comm_unique_id,
    dot_bool_batch,
    dot_u16_batch,
    LIMBS,
    ROTATIONS,
    ROWS,
    Database,
    DistanceEngine,
    IrisError,
    MasksEngine,
    combine_min,
    combine_min_batch,
    match_min,
    match_min_async,
    denominators,
    denominators_batch,
    device_count,
    distances,
    distances_batch,
    dot_bool,
    encode,
    engines_from_templates,
    dot_u16,
    launch_count,
    lib,
    library_path,
    match,
    match_streamed,
)
