"""libsvm ingest: the native reader of libps_b200.so against the Python restatement of
LibsvmParser.parse / CTR.parseFeature / DataSource.readLine / DataSet.run (oracle/libsvm_oracle.py).  Host code only: runs without a GPU."""
import gzip
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import libsvm_oracle as lo  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF_TRAIN = "/root/reference/src/main/resources/train.txt"


@pytest.fixture(scope="module")
def sample(tmp_path_factory):
    txt = gzip.open(os.path.join(GOLD, "ctr_sample.txt.gz"), "rb").read()
    p = tmp_path_factory.mktemp("ctr") / "sample.txt"
    p.write_bytes(txt)
    return str(p), txt.decode().split("\n")[:-1]


def _eq(a, b):
    for k in ("E", "W", "Y", "X"):
        assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape, k
        assert np.array_equal(a[k].view(np.uint8), b[k].view(np.uint8)), k      # bit-exact, floats included


def test_oracle_reproduces_golden(sample):
    _, lines = sample
    g = np.load(os.path.join(GOLD, "ctr_sample.npz"))
    b = lo.parse_feature([lo.parse_line(ln) for ln in lines])
    _eq(b, {k: g[k] for k in g.files})
    assert b["E"].shape == (320, 23) and np.array_equal(b["W"], b["E"] % 100000)   # ids < 2^24: float % == integer %
    assert set(np.unique(b["Y"])) <= {0.0, 1.0}


def test_native_lines_match_oracle(ps, sample):
    _, lines = sample
    for ln in lines[:64]:
        st, E, X, W, Y = ps.parse_libsvm_line(ln)
        o = lo.parse_feature([lo.parse_line(ln)])
        assert st == 0
        _eq(dict(E=E[None], X=X[None], W=W[None], Y=np.array([Y], np.float32)), o)


@pytest.mark.parametrize("line,status", [
    ("", 1), ("   ", 1), ("\t", 1),                                  # StringUtils.isBlank -> empty list -> IndexOutOfBounds later
    ("1 0:1 2:1", 1),                                                # short line
    ("1  0:1", 2),                                                   # inner empty token: Long.parseLong("")
    (" 1 0:1", 2),                                                   # leading empty token: Float.parseFloat("")
    ("x 0:1", 2), ("1 a:1", 2), ("1 5", 2), ("1 5:", 2), ("1 :5", 2), ("1 5:x", 2), ("1 5:1:", 1), ("1 9223372036854775808:1", 2),
    ("1 +5:1 ", 1), ("1.5e0 5:0x1p-1", 1), ("1f 5:2d", 1), ("NaN 5:-Infinity", 1), ("nan 5:1", 2), ("1 5:inf", 2), ("1 5:1e", 2), ("1 5:.", 2),
])
def test_malformed_lines_follow_java(ps, line, status):
    st = ps.parse_libsvm_line(line, F=23, Xn=45)[0]
    try:
        cols = lo.parse_line(line)
        lo.parse_feature([cols])
        exp = 0
    except lo.JavaException as e:
        exp = 1 if "IndexOutOfBounds" in str(e) and "pair" not in str(e) and "cols" not in str(e) else 2
    assert st == exp == status, (line, st, exp)


def test_float_spellings_bit_exact(ps):
    rng = np.random.default_rng(1)
    toks = ["0", "-0", "1", "0.5", "0.48", "1e-3", "123456789", "16777217", "0.1", "3.4028235e38", "1e-45", "7.0064923216240854e-46", "1.17549435E-38",
            "0.000001", "33554433", "9007199254740993", ".5", "5.", "+2.5", "1E2", "0x1.fffffep127", "0.3f"]
    toks += [repr(float(x)) for x in rng.standard_normal(200) * 10.0 ** rng.integers(-8, 8, 200)]
    toks += ["%.2f" % x for x in rng.random(100)]
    for t in toks:
        st, E, X, W, Y = ps.parse_libsvm_line(f"{t} 7:{t}", F=0, Xn=1)
        assert st == 0, t
        exp = lo.parse_float(t)
        assert np.float32(Y).view(np.uint32) == exp.view(np.uint32) and X[0].view(np.uint32) == exp.view(np.uint32), (t, Y, exp)


def test_ids_above_2_24_follow_the_float_cast(ps):
    st, E, X, W, Y = ps.parse_libsvm_line("0 16777217:1 123456789:1 -7:1", F=3, Xn=0)
    o = lo.parse_feature([lo.parse_line("0 16777217:1 123456789:1 -7:1")], F=3, Xn=0)
    assert st == 0 and np.array_equal(E, o["E"][0]) and np.array_equal(W, o["W"][0])
    assert E[0] == 16777216 and E[1] == 123456792 and W[2] == -7          # (float) idx; Java % keeps the dividend's sign


@pytest.mark.parametrize("batch,offset,step,threads", [(100, 0, 1, 1), (64, 0, 1, 3), (1000, 0, 1, 2), (37, 1, 2, 1), (50, 3, 4, 2), (1, 0, 1, 1), (16, 319, 1, 1), (16, 500, 1, 1)])
def test_reader_batches_match_dataset(ps, sample, batch, offset, step, threads):
    path, lines = sample
    exp = [b for b in lo.dataset_batches(lines, batch, offset=offset, step=step) if b is not None]
    r = ps.LibsvmReader(path, batch=batch, offset=offset, step=step, threads=threads)
    for epoch in range(2):                                    # DataSet.reset rewinds
        got = list(r)
        assert len(got) == len(exp)
        for a, b in zip(got, exp):
            _eq(a, b)
        assert r.next() is None and r.next() is None          # DataSet.next() stays null at end of data
        r.reset()
    r.close()


def test_reader_loses_the_batches_the_reference_loses(ps, sample, tmp_path):
    _, lines = sample
    bad = list(lines[:120])
    bad[7] = "1 oops"                   # parse exception in batch 0: lines 0..7 lost, batch restarts at line 8
    bad[40] = ""                        # blank line: the whole batch that holds it is lost in parseFeature
    bad[95] = "0 1:1 2:1"               # short line: same
    bad[96] = "0 1:1 2:x"               # ... and a parse exception right after it in the same batch
    p = tmp_path / "bad.txt"
    p.write_text("\r\n".join(bad) + "\r\n")      # CRLF terminators too
    exp_all = list(lo.dataset_batches(bad, 16))
    exp = [b for b in exp_all if b is not None]
    r = ps.LibsvmReader(str(p), batch=16, threads=2)
    got = list(r)
    assert len(got) == len(exp) and len(exp) < len(exp_all)
    for a, b in zip(got, exp):
        _eq(a, b)
    st = r.stats()
    assert st["dropped_batches"] == sum(b is None for b in exp_all) and st["batches"] == len(exp)
    r.close()


def test_reader_errors(ps, tmp_path):
    with pytest.raises(ps.PsError) as e:
        ps.LibsvmReader(str(tmp_path / "missing.txt"))
    assert e.value.code == 204
    with pytest.raises(ps.PsError):
        ps.LibsvmReader(__file__, batch=0)
    empty = tmp_path / "empty.txt"
    empty.write_text("")
    r = ps.LibsvmReader(str(empty))
    assert r.next() is None
    r.close()
    nonl = tmp_path / "nonl.txt"                               # last line without a terminator is still a line (BufferedReader)
    nonl.write_text("1 5:1\n0 6:2")
    r = ps.LibsvmReader(str(nonl), F=1, Xn=0, batch=10)
    b = r.next()
    assert b["E"].tolist() == [[5], [6]] and b["Y"].tolist() == [1.0, 0.0] and r.next() is None
    r.close()


@pytest.mark.skipif(not os.path.exists(REF_TRAIN), reason="the reference's bundled sample is only present in the build container")
def test_bundled_sample_statistics(ps):
    """SURVEY §8d anchors: 100 000 lines, 35 641 positives, per-field unique counts of the bundled train.txt."""
    r = ps.LibsvmReader(REF_TRAIN, batch=1000, threads=4)
    E, Y = [], []
    for b in r:
        E.append(b["E"].copy())
        Y.append(b["Y"].copy())
    st = r.stats()
    r.close()
    E, Y = np.concatenate(E), np.concatenate(Y)
    assert st == dict(lines=100000, batches=100, dropped_batches=0)
    assert int(Y.sum()) == 35641 and E.shape == (100000, 23)
    from ps_b200.synth import FIELD_UNIQUES
    assert [len(np.unique(E[:, j])) for j in range(23)] == FIELD_UNIQUES


def test_reader_reset_midway_and_early_close(ps, sample):
    """DataSet.reset (DataSet.java:61-67) may come at any time: the producer thread is stopped, the queue dropped, the source rewound."""
    path, lines = sample
    exp = [b for b in lo.dataset_batches(lines, 32) if b is not None]
    r = ps.LibsvmReader(path, batch=32, threads=2)
    for _ in range(3):
        got = [r.next(), r.next()]
        _eq(got[0], exp[0])
        _eq(got[1], exp[1])
        r.reset()
    got = list(r)
    assert len(got) == len(exp)
    for a, b in zip(got, exp):
        _eq(a, b)
    r.close()
    r2 = ps.LibsvmReader(path, batch=7)                       # closing with parsed batches still queued must not hang or leak the thread
    r2.next()
    r2.close()


# --------------------------------------------------------------------------- property test: native parser == Java semantics
def _expected(line, F, Xn):
    try:
        o = lo.parse_feature([lo.parse_line(line)], F=F, Xn=Xn)
        return 0, o
    except lo.JavaException as e:
        msg = str(e)
        return (1 if msg.startswith("IndexOutOfBounds") else 2), None


def test_random_lines_agree_with_the_java_semantics(ps):
    """Lines assembled from valid and invalid spellings of every token kind; status and (for status 0) every output bit must agree
    with the Python restatement of LibsvmParser.parse + CTR.parseFeature."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")
    f_ok = ["0", "1", "-1", "0.5", ".5", "5.", "-0", "00012.3400", "1e3", "1E-2", "1.5f", "2D", "NaN", "-Infinity", "0x1p3", "0x.8p1", "+7", "1234567", "12345678",
            "0.000000001", "0.00000000001", "3.4028235e38", "1e39", "1 ", "\t3", "4\t", "99999999999999999999", "0.48", "16777217", "0.1", "7.0064923216240854e-46"]
    f_bad = ["", ".", "-", "e5", "1e", "1..2", "nan", "inf", "0x1", "--1", "1f2"]
    i_ok = ["0", "7", "33895", "16777217", "123456789012345678", "9223372036854775807", "-3", "+4", "007", "99999"]
    i_bad = ["9223372036854775808", "", "x", "1.0", " 5"]
    floats = st.sampled_from(f_ok * 12 + f_bad)                      # mostly valid spellings, so that whole lines parse
    idxs = st.sampled_from(i_ok * 12 + i_bad)
    pair = st.tuples(idxs, floats, st.sampled_from([""] * 20 + [":", ":9", "::"])).map(lambda t: f"{t[0]}:{t[1]}{t[2]}") | st.sampled_from(["5", ":", "", "a:b"])
    good_pair = st.tuples(st.sampled_from(i_ok), st.sampled_from(f_ok)).map(lambda t: f"{t[0]}:{t[1]}")
    sep = st.sampled_from([" "] * 30 + ["  "])
    tok = st.one_of(*([good_pair] * 15 + [pair]))
    tokens = st.lists(tok, min_size=4, max_size=7) | st.lists(tok, min_size=0, max_size=7)     # F + Xn = 4 columns make a full line
    line = st.tuples(floats, tokens, st.lists(sep, min_size=7, max_size=7), st.sampled_from(["", "", " ", "   "])).map(
        lambda t: t[0] + "".join(t[2][i] + p for i, p in enumerate(t[1])) + t[3])
    seen = {0: 0, 1: 0, 2: 0}

    @hyp.settings(max_examples=3000, deadline=None, derandomize=True)
    @hyp.given(line)
    def check(ln):
        if "\n" in ln or "\r" in ln:
            return
        got = ps.parse_libsvm_line(ln, F=2, Xn=2)
        exp_status, o = _expected(ln, 2, 2)
        seen[exp_status] += 1
        assert got[0] == exp_status, (ln, got[0], exp_status)
        if exp_status == 0:
            assert np.array_equal(got[1], o["E"][0]) and np.array_equal(got[3], o["W"][0]), ln
            assert np.array_equal(got[2].view(np.uint32), o["X"][0].view(np.uint32)) or (np.isnan(got[2]) == np.isnan(o["X"][0])).all(), ln
            assert np.float32(got[4]).view(np.uint32) == o["Y"][0].view(np.uint32) or (np.isnan(got[4]) and np.isnan(o["Y"][0])), ln
    check()
    assert min(seen.values()) >= 100, seen                            # all three outcomes are well represented


def test_reader_with_random_damage_matches_dataset(ps, sample, tmp_path):
    """The reader gathers the next batch while its helpers parse the current one; a line that throws inside the parser takes that lookahead
    back and restarts the batch behind it (DataSet.java:84-98).  Random damage, batch sizes, thread counts, offset / step: delivered batches,
    dropped batches and line counts must equal the Java-semantics restatement's."""
    _, lines = sample
    rng = np.random.default_rng(11)
    for trial in range(12):
        n = min(len(lines), int(rng.integers(60, 700)))
        ls = list(lines[:n])
        for _ in range(int(rng.integers(0, 9))):
            i = int(rng.integers(0, n))
            ls[i] = ["1 oops", "", "0 1:1 2:1", "0 1:1 2:x", "   ", "1 5:"][int(rng.integers(0, 6))]
        p = tmp_path / f"damaged{trial}.txt"
        p.write_text("\n".join(ls) + ("\n" if rng.random() < 0.7 else ""))
        batch = int(rng.choice([1, 3, 16, 64, 65, 130, 500]))
        offset, step = (int(rng.integers(0, 5)), int(rng.integers(1, 4))) if rng.random() < 0.4 else (0, 1)
        threads = int(rng.integers(1, 6))
        exp_all = list(lo.dataset_batches(ls, batch, offset=offset, step=step))
        exp = [b for b in exp_all if b is not None]
        r = ps.LibsvmReader(str(p), batch=batch, offset=offset, step=step, threads=threads)
        for epoch in range(2):
            got = list(r)
            assert len(got) == len(exp), (trial, batch, offset, step, threads)
            for a, b in zip(got, exp):
                _eq(a, b)
            r.reset()
        r.close()
