"""Cuts the small libsvm fixture the ingest tests travel with from the reference's bundled sample and records what the CPU
restatement (oracle/libsvm_oracle.py) makes of it.

  python scripts/make_golden_ctr.py            # needs /root/reference (this container only)

tests/golden/ctr_sample.txt.gz   = the first 256 lines of src/main/resources/train.txt + the first 64 of test.txt
tests/golden/ctr_sample.npz      = E, X, W, Y of those 320 lines through parse_line + parse_feature (CTR.java:47-68)
"""
import gzip
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import libsvm_oracle as lo  # noqa: E402

RES = "/root/reference/src/main/resources"
lines = lo.read_lines(os.path.join(RES, "train.txt"))[:256] + lo.read_lines(os.path.join(RES, "test.txt"))[:64]
out = os.path.join(ROOT, "tests", "golden")
with gzip.GzipFile(os.path.join(out, "ctr_sample.txt.gz"), "wb", mtime=0) as f:
    f.write(("\n".join(lines) + "\n").encode())
b = lo.parse_feature([lo.parse_line(ln) for ln in lines])
np.savez_compressed(os.path.join(out, "ctr_sample.npz"), **b)
print({k: v.shape for k, v in b.items()}, "label mean", float(b["Y"].mean()))
