"""CPU restatement of the reference's libsvm ingest — TEST INFRASTRUCTURE ONLY (tests/ and bench cpu_baseline may import it;
nothing under ps_b200/ does).  Pure Python, line by line after the Java (paths relative to /root/reference/src/main/java/):

  parse_line      data/LibsvmParser.java:13-25   (String.split(" "), split(":"), Long.parseLong, Float.parseFloat)
  parse_feature   CTR.java:47-68                 (Y, E = (float) idx, X = value, W = MatrixUtil.hash(E, 100000) MatrixUtil.java:27-33)
  DataSource      data/DataSource.java:25-46     (offset / step line selection)
  dataset_batches data/DataSet.java:77-100       (batching; exceptions swallowed by `catch (Exception e) { // ignore }`)

Parity unpinned by the reference itself (no JVM here; the reference's tests hold no fixture for this path): pinned by the
properties tests/test_ingest.py states and by the bundled sample (100 000 lines, label ratio 35 641 / 100 000 — SURVEY 8d).
"""
import re

import numpy as np

_LONG = re.compile(r"[+-]?[0-9]+\Z")
_FLOAT = re.compile(r"[+-]?(NaN|Infinity|(([0-9]+\.?[0-9]*|\.[0-9]+)([eE][+-]?[0-9]+)?|0[xX]([0-9a-fA-F]+\.?[0-9a-fA-F]*|\.[0-9a-fA-F]+)[pP][+-]?[0-9]+)[fFdD]?)\Z")


class JavaException(Exception):
    pass


def java_split(s, sep):
    """String.split(single-char literal): trailing empty strings removed; "" -> [""]."""
    parts = s.split(sep)
    while len(parts) > 1 and parts[-1] == "":
        parts.pop()
    if parts == [""] and s != "":
        return []            # e.g. " ".split(" ") -> [] in Java
    return parts


def parse_long(s):
    if not _LONG.match(s):
        raise JavaException("NumberFormatException: " + s)
    v = int(s)
    if not -(1 << 63) <= v < (1 << 63):
        raise JavaException("NumberFormatException: " + s)
    return v


def parse_float(s):
    s = s.strip("".join(chr(c) for c in range(33)))          # String.trim(): chars <= ' '
    if not _FLOAT.match(s):
        raise JavaException("NumberFormatException: " + s)
    body = s.rstrip("fFdD") if not s.lower().lstrip("+-").startswith("0x") else s[:-1] if s[-1] in "fFdD" else s
    if body.lstrip("+-") == "NaN":
        return np.float32(np.nan)
    if body.lstrip("+-") == "Infinity":
        return np.float32(-np.inf if body[0] == "-" else np.inf)
    if body.lower().lstrip("+-").startswith("0x"):
        return np.float32(float.fromhex(body))               # double then float: hex literals this short round once
    return nearest_float32(body)


def nearest_float32(dec):
    """Correctly rounded (nearest, ties to even) float32 of a decimal literal, as Float.parseFloat returns — computed exactly
    with rationals so that no intermediate double rounding can leak in."""
    from fractions import Fraction
    exact = Fraction(dec)
    with np.errstate(over="ignore"):
        f = np.float32(float(dec))
    if not np.isfinite(f) or exact == 0:
        return f
    best = None
    with np.errstate(over="ignore"):
        cands = (np.nextafter(f, np.float32(-np.inf)), f, np.nextafter(f, np.float32(np.inf)))
    for cand in cands:
        if not np.isfinite(cand):
            continue
        err = abs(Fraction(float(cand)) - exact)
        even = (int(np.float32(cand).view(np.uint32)) & 1) == 0
        key = (err, 0 if even else 1)
        if best is None or key < best[0]:
            best = (key, cand)
    return np.float32(best[1])


def is_blank(s):
    return all(ch.isspace() for ch in s)


def parse_line(line):
    """LibsvmParser.parse: list of (idx, value); [] for a blank line."""
    if is_blank(line):
        return []
    cols = java_split(line, " ")
    if not cols:
        raise JavaException("ArrayIndexOutOfBounds: cols[0]")
    out = [(0, parse_float(cols[0]))]
    for c in cols[1:]:
        pair = java_split(c, ":")
        if len(pair) < 2:
            if len(pair) == 1:
                parse_long(pair[0])                           # evaluated first; may throw NumberFormatException instead
            raise JavaException("ArrayIndexOutOfBounds: pair[1]")
        out.append((parse_long(pair[0]), parse_float(pair[1])))
    return out


def parse_feature(data_list, F=23, Xn=45, wide_size=100000):
    """CTR.parseFeature: dict E (N,F) int64 [= (long)(float) idx], X (N,Xn) f32, W (N,F) int64, Y (N,) f32."""
    N = len(data_list)
    E = np.zeros((N, F), np.float32)
    X = np.zeros((N, Xn), np.float32)
    Y = np.zeros(N, np.float32)
    for i, cols in enumerate(data_list):
        if len(cols) < 1 + F + Xn:
            raise JavaException("IndexOutOfBoundsException")
        Y[i] = cols[0][1]
        for j in range(1, 1 + F):
            E[i, j - 1] = np.float32(cols[j][0])              # long -> float
        for j in range(1 + F, 1 + F + Xn):
            X[i, j - 1 - F] = cols[j][1]
    W = np.fmod(E, np.float32(wide_size))                     # Java float % == C fmodf
    with np.errstate(invalid="ignore"):                       # ids beyond the long range: Java's (long) cast saturates; out of the tested domain
        return dict(E=E.astype(np.int64), X=X, W=W.astype(np.int64), Y=Y)


class DataSource:
    """data/DataSource.java:25-46 over a list of lines."""

    def __init__(self, lines, offset=0, step=1):
        self.lines, self.offset, self.step = lines, offset, step
        self.reset()

    def reset(self):
        self.idx, self.cursor = 0, 0

    def _read_internal(self):
        if self.cursor >= len(self.lines):
            return None
        s = self.lines[self.cursor]
        self.cursor += 1
        return s

    def read_line(self):
        while self.idx <= self.offset:
            line = self._read_internal()
            self.idx += 1
            if self.idx - 1 == self.offset:
                return line
        line = None
        for _ in range(self.step):
            line = self._read_internal()
            self.idx += 1
        return line


def dataset_batches(lines, batch, F=23, Xn=45, wide_size=100000, offset=0, step=1):
    """DataSet.run with one reader thread: yields the batches that reach the queue; returns via StopIteration at eof.
    Also counts the batches the swallowed exceptions lose (attribute .dropped on the generator's frame is awkward: use stats list)."""
    src = DataSource(lines, offset, step)
    eof = False
    while not eof:
        try:
            data_list = []
            for _ in range(batch):
                line = src.read_line()
                if line is None:
                    eof = True
                    break
                data_list.append(parse_line(line))
            if not data_list:
                continue
            yield parse_feature(data_list, F, Xn, wide_size)
        except JavaException:
            yield None                                        # a lost batch (the consumer never sees it)


def read_lines(path):
    """BufferedReader.readLine semantics for \\n and \\r\\n terminated files."""
    with open(path, "rb") as f:
        data = f.read().decode("latin-1")
    lines = data.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    return [ln[:-1] if ln.endswith("\r") else ln for ln in lines]
