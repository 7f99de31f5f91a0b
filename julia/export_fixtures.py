"""Replays the seeded theta generator of the golden fixtures whose parameter vector is stored as a digest only (the
1.3 M-parameter RealNVP d=64 case) and writes it as raw little-endian Float32, so that julia/crosscheck.jl can cover the
headline configuration on a machine that has Julia:

    python julia/export_fixtures.py /tmp/nf_fixtures
    julia --project=<NormalizingFlows.jl>/test julia/crosscheck.jl tests/golden /tmp/nf_fixtures
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
from helpers import oracle_flow  # noqa: E402


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    gdir = os.path.join(ROOT, "tests", "golden")
    for fn in sorted(os.listdir(gdir)):
        if not fn.endswith(".npz"):
            continue
        d = np.load(os.path.join(gdir, fn), allow_pickle=False)
        if "theta" in d.files:
            continue
        meta = json.loads(str(d["meta"]))
        theta32 = oracle_flow(meta["kind"], meta["dim"], np.float64, **meta["kw"]).theta().numpy().astype(np.float32)
        assert hashlib.sha256(theta32.tobytes()).hexdigest() == str(d["theta_sha256"]), fn
        path = os.path.join(out_dir, fn[:-4] + ".theta.f32")
        theta32.tofile(path)
        print(path, theta32.size, "float32 values")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "nf_fixtures")
