#!/usr/bin/env python
"""Pack the raw arrays written by tools/dump_reference_vectors.jl (run with the real ClimaAtmos stack on a machine that has julia)
into the fixtures tests/test_reference_fixtures.py consumes:  python tools/ref_vectors_to_npz.py <dump dir> tests/golden

Julia arrays are column-major; reading the bytes with the REVERSED shape gives the C-order view used everywhere in this repo:
a VIJFH parent array (Nv, Nq, Nq, Nf, Nh) becomes [h, f, j, i, v]."""
import json
import os
import sys

import numpy as np

DT = {"Float32": "<f4", "Float64": "<f8", "Int32": "<i4", "Int64": "<i8", "UInt8": "u1"}


def main(src, dst):
    idx = json.load(open(os.path.join(src, "index.json")))
    cases = {}
    for key, meta in idx.items():
        case, name = key.split("/", 1)
        a = np.fromfile(os.path.join(src, meta["file"]), dtype=DT[meta["dtype"]]).reshape(tuple(reversed(meta["shape"])))
        cases.setdefault(case, {})[name] = a
    os.makedirs(dst, exist_ok=True)
    for case, arrays in cases.items():
        out = os.path.join(dst, f"ref_{case}.npz")
        np.savez_compressed(out, **arrays)
        print(out, len(arrays), "arrays", round(os.path.getsize(out) / 1e6, 2), "MB")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
