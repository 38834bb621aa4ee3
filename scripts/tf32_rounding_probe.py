#!/usr/bin/env python
"""Does tcgen05.mma kind::tf32 TRUNCATE or ROUND its fp32 operands to TF32?  (decides whether the 3xTF32 split has to
write the explicit high part back).  x = 1 + 3*2^-12 lies between the TF32 neighbours 1 and 1 + 2^-10:
truncation gives 1, round-to-nearest gives 1 + 2^-10."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ps_b200 import binding as ps  # noqa: E402

ctx = ps.Context(0, seed=1)
for x in (1.0 + 3 * 2.0 ** -12, 1.0 + 2.0 ** -11 + 2.0 ** -20, -(1.0 + 3 * 2.0 ** -12), 1.0 + 2.0 ** -11):
    A = np.zeros((128, 32), np.float32)
    B = np.zeros((64, 32), np.float32)
    A[:, 0] = x
    B[:, 0] = 1.0
    c = ctx.gemm_nt(ps.PS_FC_TF32, A, B)
    print(f"x={x!r} -> a*1 = {float(c[0, 0])!r}  (trunc={float(np.float32((np.float32(x).view(np.uint32) & 0xFFFFE000).view(np.float32)))!r})")
