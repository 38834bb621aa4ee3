"""GPU-side libsvm parse (ps_libsvm_parse_dev) against the host parser (ps_libsvm_parse_line / the reader), which tests/test_ingest.py
pins against the Python restatement of LibsvmParser.parse + CTR.parseFeature."""
import gzip
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def parse_dev(ps, ctx, text, F, Xn, max_rows, wide=100000):
    t = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
    E = torch.full((max_rows, max(F, 1)), -7, dtype=torch.int64, device="cuda")
    W = torch.full((max_rows, max(F, 1)), -7, dtype=torch.int64, device="cuda")
    X = torch.full((max_rows, max(Xn, 1)), -7.0, dtype=torch.float32, device="cuda")
    Y = torch.full((max_rows,), -7.0, dtype=torch.float32, device="cuda")
    st = torch.full((max_rows,), 9, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    rows = ctx.parse_libsvm_dev(t.data_ptr(), len(text), F, Xn, wide, max_rows, E.data_ptr(), X.data_ptr(), W.data_ptr(), Y.data_ptr(), st.data_ptr())
    ctx.synchronize()
    return rows, E.cpu().numpy()[:, :F], X.cpu().numpy()[:, :Xn], W.cpu().numpy()[:, :F], Y.cpu().numpy(), st.cpu().numpy()


def host_rows(ps, lines, F, Xn):
    out = [ps.parse_libsvm_line(ln, F=F, Xn=Xn) for ln in lines]
    return (np.array([o[0] for o in out]), np.stack([o[1] for o in out]), np.stack([o[2] for o in out]), np.stack([o[3] for o in out]),
            np.array([o[4] for o in out], np.float32))


def test_fixture_lines_bit_identical_to_host(ps, ctx):
    text = gzip.open(os.path.join(GOLD, "ctr_sample.txt.gz"), "rb").read()
    lines = text.decode().split("\n")[:-1]
    rows, E, X, W, Y, st = parse_dev(ps, ctx, text, 23, 45, 512)
    assert rows == len(lines) == 320 and (st[:rows] == 0).all() and (st[rows:] == 9).all()
    g = np.load(os.path.join(GOLD, "ctr_sample.npz"))
    assert np.array_equal(E[:rows], g["E"]) and np.array_equal(W[:rows], g["W"])
    assert np.array_equal(X[:rows].view(np.uint32), g["X"].view(np.uint32)) and np.array_equal(Y[:rows].view(np.uint32), g["Y"].view(np.uint32))


def test_statuses_and_deferral_to_host(ps, ctx):
    good = "1 7:1 8:0.25 9:0.5"
    lines = [good, "", "0 7:1", "0 7:1 8:1e-3 9:2", "0  7:1 8:1 9:1", "0 7:1:3 8:1 9:1", "1 16777217:1 8:-0.125 9:.5", good + "   ", "0 7:1 8:1 9:1 10:1 11:1",
             "x 7:1 8:1 9:1", "0 -7:1 8:1 9:1", "0 7:1 8:0.12345678 9:1", "0 123456789012345678:1 8:1 9:1"]
    text = ("\r\n".join(lines) + "\r\n").encode()                      # CRLF terminators
    rows, E, X, W, Y, st = parse_dev(ps, ctx, text, 1, 2, 64)
    assert rows == len(lines)
    hst, hE, hX, hW, hY = host_rows(ps, lines, 1, 2)
    assert st[:rows].tolist() == [0, 1, 1, 2, 2, 2, 0, 0, 0, 2, 2, 2, 0]
    for r in range(rows):
        if st[r] == 2:
            continue                                                  # deferred: the host decides (ok, or the batch-dropping kinds)
        assert st[r] == hst[r], (r, lines[r])
        if st[r] == 0:
            assert np.array_equal(E[r], hE[r]) and np.array_equal(W[r], hW[r]), (r, lines[r])
            assert np.array_equal(X[r].view(np.uint32), hX[r].view(np.uint32)) and np.float32(Y[r]).view(np.uint32) == hY[r].view(np.uint32), (r, lines[r])
    assert E[6, 0] == 16777216 and W[12, 0] == hW[12, 0]              # (float) idx above 2^24; float % for huge ids
    assert hst[3] == 0 and hst[5] == 0 and hst[4] == 2                # what the host makes of three of the deferred lines


def test_many_lines_cross_scan_tiles(ps, ctx):
    """20 000 synthetic CTR lines (13.9 MB: more than 1024 4-KB chunks, so the chunk scan walks several tiles), no trailing newline on
    the last line: it is not a complete line and is left out."""
    from ps_b200.synth import Synth
    syn = Synth(F=23, Xn=45, V=1_000_000, seed=41)
    b = syn.batch(20000)
    lines = []
    for n in range(20000):
        cols = ["%d" % int(b["Y"][n])] + ["%d:1" % int(b["E"][n, j]) for j in range(23)] + ["%d:%.2f" % (33895 + x, b["X"][n, x]) for x in range(45)]
        lines.append(" ".join(cols))
    text = "\n".join(lines).encode()
    rows, E, X, W, Y, st = parse_dev(ps, ctx, text, 23, 45, 20000)
    assert rows == 19999 and (st[:rows] == 0).all()
    assert np.array_equal(E[:rows], b["E"][:rows]) and np.array_equal(W[:rows], b["W"][:rows])
    assert np.array_equal(X[:rows].view(np.uint32), b["X"][:rows].view(np.uint32)) and np.array_equal(Y[:rows], b["Y"][:rows])
    rows2, E2, *_ = parse_dev(ps, ctx, text + b"\n", 23, 45, 5000)     # more lines than max_rows: the first max_rows are parsed
    assert rows2 == 5000 and np.array_equal(E2[:5000], b["E"][:5000])
