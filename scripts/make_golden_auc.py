"""Extracts the hard-coded (p, y) vector of the reference's T/TestAuc.java (the only numeric fixture its
tests hold) and records the AUC that evaluate/AUC.java:32-82 computes on it, evaluated here by a plain
Python transcription of that method (stable sort by p, sweep from the top, rectangle rule).  Run in the
build container only (reads /root/reference); the .npz it writes is the committed fixture."""
import re
import sys

import numpy as np

src = open("/root/reference/src/test/java/TestAuc.java").read()
strings = re.findall(r'"([0-9eE+\-., ]{200,})"', src)
assert len(strings) == 2, len(strings)
p = np.array([float(t) for t in strings[0].split(",")])
y = np.array([float(t) for t in strings[1].split(",")])
assert p.shape == y.shape

order = np.argsort(p, kind="stable")
pos = float((y > 0).sum())
neg = float(len(y) - pos)
tp = fp = 0.0
prev = auc = 0.0
for i in order[::-1]:
    if y[i] > 0.0:
        fp += 1
    else:
        tp += 1
    x, yy = tp / pos, fp / neg
    if x != prev:
        auc += (x - prev) * yy
        prev = x
print(len(p), auc)
np.savez_compressed(sys.argv[1] if len(sys.argv) > 1 else "tests/golden/testauc_vector.npz", p=p.astype(np.float32), y=y.astype(np.float32), auc=np.float64(auc))
