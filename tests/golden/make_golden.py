"""Regenerates tests/golden/golden_small.json.

The reference (Rust) cannot be built in this image, so these vectors come from the numpy
DEFINITION in oracle/ (np_* functions: np.roll rotation + integer matmul), not from the C
restatement they are used to check.  Inputs are re-derivable from seeds via the synthetic
row generator, so only seeds and outputs are stored.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle as O  # noqa: E402

SEED, ROW0, N = 0x1715C0DE, 1000, 24


def main():
    db = O.gen_share_rows(SEED, ROW0, N)
    masks = O.gen_mask_rows(SEED, ROW0, N)
    cases = []
    for kind, qseed in (("ternary", 11), ("ternary", 12), ("uniform", 21)):
        qm = O.gen_mask_rows(qseed, 1, 1)[0]
        if kind == "ternary":
            q = O.np_encode(O.gen_mask_rows(qseed, 0, 1)[0], qm)
        else:
            q = O.gen_share_rows(qseed, 0, 1)[0]
        cases.append(
            {
                "kind": kind,
                "qseed": qseed,
                "distances": O.np_distance_batch(q, db).tolist(),
                "denominators": O.np_masks_batch(qm, masks).tolist(),
            }
        )
    i = np.arange(O.BITS)
    secret = (((i // O.COLS) << 8) | (i % O.COLS)).astype(np.uint16)
    # reference known-answer (src/encoded_bits.rs:205-219): rotated(a)[row][col] = row<<8 | (col-a) mod 200
    kn = {str(a): (((i[:32] // O.COLS) << 8) | ((O.COLS + i[:32] % O.COLS - a) % O.COLS)).tolist() for a in (-15, -1, 0, 1, 15)}
    assert all(O.np_encoded_rotated(secret, int(a))[:32].tolist() == v for a, v in kn.items())
    out = {"seed": SEED, "row0": ROW0, "n_rows": N, "cases": cases, "rotated_number": kn}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_small.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
