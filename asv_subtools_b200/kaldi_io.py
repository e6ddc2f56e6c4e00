"""Kaldi ark/scp reader-writer for the extraction / scoring CLIs.

Byte-compatible with the formats the reference's `pytorch/libs/support/kaldi_io.py` speaks
(`open_or_fd` :43-73, `read_key` :148-163, `read_mat` :449-569, `read_vec_flt` :329-363,
`write_vec_flt` :367-399, `write_mat` :573-608); golden byte streams produced by that module are
replayed in tests/test_kaldi_io.py.  Formats:

  ark entry   : <key> ' ' then either binary ('\\0B' + payload) or ascii
  float matrix: 'FM ' | 'DM '  + '\\4' int32 rows + '\\4' int32 cols + row-major data
  compressed  : 'CM '  + {min f32, range f32, rows i32, cols i32} + cols x 4 uint16 percentiles +
                cols x rows uint8 (column-major); value = piecewise-linear in the percentiles
  float vector: 'FV ' | 'DV '  + '\\4' int32 dim + data
  rspecifier  : [ark|scp][,opts]:<file>[:offset] | '<command> |' (input pipe) | '| <command>' (output pipe)
"""
import gzip
import os
import re
import struct
import subprocess
import sys
import threading

import numpy as np


class KaldiFormatError(Exception):
    pass


_PREFIX = re.compile(r"^(ark|scp)(,scp|,b|,t|,n?f|,n?p|,b?o|,n?s|,n?cs)*:")


class _PipeStream:
    """One end of a shell pipe that owns its child.  close() (or leaving a `with` block) waits for the
    command and raises on a nonzero exit status, so a writer such as `ark:| copy-vector ark:- ark,scp:x.ark,x.scp`
    has finished x.scp when the extractor returns and a failed pipe command fails the job (the reference's
    popen, kaldi_io.py:75-110, keeps the interpreter alive on non-daemon cleanup threads; a stream that is
    never closed is reaped the same way here)."""

    def __init__(self, cmd, mode):
        self.cmd, self.mode = cmd, mode
        if mode == "rb":
            self.proc = subprocess.Popen(cmd, shell=True, stdout=subprocess.PIPE, stderr=sys.stderr)
            self.raw = self.proc.stdout
        else:
            self.proc = subprocess.Popen(cmd, shell=True, stdin=subprocess.PIPE, stderr=sys.stderr)
            self.raw = self.proc.stdin
        self._closed = False
        # non-daemon: the interpreter does not exit before the child has (reference behaviour)
        self._reaper = threading.Thread(target=self._reap, daemon=False)
        self._reaper.start()

    def _reap(self):
        rc = self.proc.wait()
        if rc != 0 and not self._closed:
            sys.stderr.write("ERROR: command `{}` exited with {}\n".format(self.cmd, rc))

    def read(self, *a):
        return self.raw.read(*a)

    def readline(self, *a):
        return self.raw.readline(*a)

    def write(self, b):
        return self.raw.write(b)

    def flush(self):
        return self.raw.flush()

    def fileno(self):
        return self.raw.fileno()

    def tell(self):
        return self.raw.tell()

    def seek(self, *a):
        return self.raw.seek(*a)

    def seekable(self):
        return False

    def __iter__(self):
        return iter(self.raw)

    @property
    def closed(self):
        return self._closed

    def close(self):
        if self._closed:
            return
        self._closed = True
        try:
            self.raw.close()
        except BrokenPipeError:
            pass
        rc = self.proc.wait()
        self._reaper.join()
        if rc != 0 and not (self.mode == "rb" and rc in (-13, 141)):   # a reader may stop before the producer has (SIGPIPE)
            raise subprocess.CalledProcessError(rc, self.cmd)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def _popen(cmd, mode):
    return _PipeStream(cmd, mode)


def open_or_fd(spec, mode="rb"):
    """Open a file, gzipped file or pipe; pass through an already-open descriptor.  Understands the
    'ark:'/'scp:' prefixes and a trailing ':offset'."""
    if not isinstance(spec, str):
        return spec
    offset = None
    if _PREFIX.search(spec):
        spec = spec.split(":", 1)[1]
    if re.search(r":[0-9]+$", spec):
        spec, off = spec.rsplit(":", 1)
        offset = int(off)
    spec = spec.strip()
    if spec.endswith("|"):
        fd = _popen(spec[:-1], "rb")
    elif spec.startswith("|"):
        fd = _popen(spec[1:], "wb")
    elif spec == "-":
        fd = sys.stdin.buffer if "r" in mode else sys.stdout.buffer
    elif spec.endswith(".gz"):
        fd = gzip.open(spec, mode)
    else:
        fd = open(spec, mode)
    if offset is not None:
        fd.seek(offset)
    return fd


def read_key(fd):
    """Next utterance key of an ark stream, or None at end of stream."""
    chars = []
    while True:
        ch = fd.read(1)
        if ch == b"" or ch == b" ":
            break
        chars.append(ch)
    key = b"".join(chars).decode("latin1").strip()
    if key == "":
        return None
    if re.match(r"^\S+$", key) is None:
        raise KaldiFormatError("malformed key {!r}".format(key))
    return key


def _read_exact(fd, n):
    buf = fd.read(n)
    if len(buf) != n:
        raise KaldiFormatError("unexpected end of stream ({} of {} bytes)".format(len(buf), n))
    return buf


def _read_dim(fd):
    if _read_exact(fd, 1) != b"\4":
        raise KaldiFormatError("expected int32 size marker")
    return struct.unpack("<i", _read_exact(fd, 4))[0]


# ------------------------------------------------------------------ matrices
def _decode_compressed(fd):
    gmin, grange, rows, cols = struct.unpack("<ffii", _read_exact(fd, 16))
    perc = np.frombuffer(_read_exact(fd, cols * 8), dtype="<u2").reshape(cols, 4).astype(np.float32)
    perc = (perc * np.float32(grange) * np.float32(1.52590218966964e-05) + np.float32(gmin)).astype(np.float32)
    data = np.frombuffer(_read_exact(fd, cols * rows), dtype=np.uint8).reshape(cols, rows)
    p0, p25, p75, p100 = (perc[:, i:i + 1] for i in range(4))
    d = data.astype(np.float32)
    lo = p0 + (p25 - p0) / np.float32(64.0) * d
    mid = p25 + (p75 - p25) / np.float32(128.0) * (d - 64)
    hi = p75 + (p100 - p75) / np.float32(63.0) * (d - 192)
    out = np.where(data <= 64, lo, np.where(data > 192, hi, mid)).astype(np.float32)
    return np.ascontiguousarray(out.T)


def _read_mat_binary(fd):
    header = _read_exact(fd, 3).decode("latin1")
    if header == "CM ":
        return _decode_compressed(fd)
    if header.startswith("CM"):
        raise KaldiFormatError("compressed format {!r} is not supported (CM2/CM3)".format(header))
    if header == "FM ":
        dtype = "<f4"
    elif header == "DM ":
        dtype = "<f8"
    else:
        raise KaldiFormatError("unknown matrix header {!r}".format(header))
    rows, cols = _read_dim(fd), _read_dim(fd)
    data = np.frombuffer(_read_exact(fd, rows * cols * np.dtype(dtype).itemsize), dtype=dtype)
    return data.reshape(rows, cols)


def _read_mat_ascii(fd):
    rows = []
    while True:
        line = fd.readline().decode("latin1")
        if not line:
            raise KaldiFormatError("end of stream inside an ascii matrix")
        toks = line.split()
        if not toks:
            continue
        last = toks[-1] == "]"
        if last:
            toks = toks[:-1]
        if toks:
            rows.append(np.array(toks, dtype=np.float32))
        if last:
            return np.vstack(rows) if rows else np.zeros((0, 0), dtype=np.float32)


def read_mat(file_or_fd):
    """One Kaldi matrix (binary FM/DM/CM or ascii) -> 2-D ndarray."""
    fd = open_or_fd(file_or_fd)
    try:
        flag = _read_exact(fd, 2)
        if flag == b"\0B":
            return _read_mat_binary(fd)
        if flag == b" [":
            return _read_mat_ascii(fd)
        raise KaldiFormatError("neither binary nor ascii matrix start: {!r}".format(flag))
    finally:
        if fd is not file_or_fd:
            fd.close()


def read_mat_ark(file_or_fd):
    fd = open_or_fd(file_or_fd)
    try:
        while True:
            key = read_key(fd)
            if key is None:
                return
            yield key, read_mat(fd)
    finally:
        if fd is not file_or_fd:
            fd.close()


def read_mat_scp(file_or_fd):
    fd = open_or_fd(file_or_fd)
    try:
        for line in fd:
            key, rx = line.decode("latin1").strip().split(None, 1)
            yield key, read_mat(rx)
    finally:
        if fd is not file_or_fd:
            fd.close()


def write_mat(file_or_fd, m, key=""):
    if not (isinstance(m, np.ndarray) and m.ndim == 2):
        raise TypeError("write_mat expects a 2-D ndarray")
    fd = open_or_fd(file_or_fd, "wb")
    try:
        if key:
            fd.write((key + " ").encode("latin1"))
        fd.write(b"\0B")
        if m.dtype == np.float32:
            fd.write(b"FM ")
        elif m.dtype == np.float64:
            fd.write(b"DM ")
        else:
            raise TypeError("unsupported dtype {}".format(m.dtype))
        fd.write(b"\4" + struct.pack("<I", m.shape[0]) + b"\4" + struct.pack("<I", m.shape[1]))
        fd.write(np.ascontiguousarray(m).tobytes())
    finally:
        if fd is not file_or_fd:
            fd.close()


# ------------------------------------------------------------------ vectors
def read_vec_flt(file_or_fd):
    fd = open_or_fd(file_or_fd)
    try:
        flag = _read_exact(fd, 2)
        if flag == b"\0B":
            header = _read_exact(fd, 3).decode("latin1")
            if header == "FV ":
                dtype = "<f4"
            elif header == "DV ":
                dtype = "<f8"
            else:
                raise KaldiFormatError("unknown vector header {!r}".format(header))
            dim = _read_dim(fd)
            if dim == 0:
                return np.array([], dtype=np.float32)
            return np.frombuffer(_read_exact(fd, dim * np.dtype(dtype).itemsize), dtype=dtype)
        toks = (flag + fd.readline()).decode("latin1").split()
        return np.array([t for t in toks if t not in ("[", "]")], dtype=float)
    finally:
        if fd is not file_or_fd:
            fd.close()


def read_vec_flt_ark(file_or_fd):
    fd = open_or_fd(file_or_fd)
    try:
        while True:
            key = read_key(fd)
            if key is None:
                return
            yield key, read_vec_flt(fd)
    finally:
        if fd is not file_or_fd:
            fd.close()


def read_vec_flt_scp(file_or_fd):
    fd = open_or_fd(file_or_fd)
    try:
        for line in fd:
            key, rx = line.decode("latin1").strip().split(None, 1)
            yield key, read_vec_flt(rx)
    finally:
        if fd is not file_or_fd:
            fd.close()


def read_vectors(spec):
    """'ark:...' / 'scp:...' / bare path (by extension) -> generator of (key, vector)."""
    if isinstance(spec, str) and (spec.startswith("scp") or spec.endswith(".scp")):
        return read_vec_flt_scp(spec)
    return read_vec_flt_ark(spec)


def write_vec_flt(file_or_fd, v, key=""):
    if not isinstance(v, np.ndarray):
        raise TypeError("write_vec_flt expects an ndarray")
    v = v.reshape(-1)
    fd = open_or_fd(file_or_fd, "wb")
    try:
        if key:
            fd.write((key + " ").encode("latin1"))
        fd.write(b"\0B")
        if v.dtype == np.float32:
            fd.write(b"FV ")
        elif v.dtype == np.float64:
            fd.write(b"DV ")
        else:
            raise TypeError("unsupported dtype {}".format(v.dtype))
        fd.write(b"\4" + struct.pack("<I", v.shape[0]))
        fd.write(np.ascontiguousarray(v).tobytes())
    finally:
        if fd is not file_or_fd:
            fd.close()


def write_vec_ark_scp(ark_path, scp_path, items):
    """Write (key, vector) pairs to an ark file plus the matching scp (what `copy-vector ark:-
    ark,scp:...` produces at the end of the reference's wspecifier, extract_xvectors_for_pytorch.sh:120)."""
    with open(ark_path, "wb") as ark, open(scp_path, "w") as scp:
        for key, vec in items:
            ark.write((key + " ").encode("latin1"))
            scp.write("{} {}:{}\n".format(key, os.path.abspath(ark_path), ark.tell()))
            write_vec_flt(ark, np.asarray(vec), key="")


# ------------------------------------------------------------------ native reader / writer
def read_mat_ark_native(rspecifier):
    """Same stream of (key, float32 matrix) as read_mat_ark / read_mat_scp, decoded by the C library
    (csrc/ark_io.cpp: whole-matrix freads instead of the reference's byte-at-a-time key loop,
    kaldi_io.py:148-163).  'DM ' matrices come back as float64 like from read_mat (values rounded through
    float32 by the C reader)."""
    import ctypes as C
    from ._lib import lib, last_error
    h = C.c_void_p()
    if lib.xvb_ark_reader_open(C.byref(h), rspecifier.encode()) != 0:
        raise KaldiFormatError(last_error())
    try:
        key, rows, cols, data = C.c_char_p(), C.c_int(), C.c_int(), C.POINTER(C.c_float)()
        while True:
            rc = lib.xvb_ark_reader_next(h, C.byref(key), C.byref(rows), C.byref(cols), C.byref(data))
            if rc == 0:
                return
            if rc < 0:
                raise KaldiFormatError(last_error())
            n = rows.value * cols.value
            m = np.ctypeslib.as_array(data, shape=(n,)).copy().reshape(rows.value, cols.value) if n else \
                np.zeros((rows.value, cols.value), dtype=np.float32)
            yield key.value.decode("latin1"), (m.astype(np.float64) if rc == 2 else m)
    finally:
        lib.xvb_ark_reader_close(h)


class NativeVectorWriter:
    """write_vec_flt through the C library: 'ark:f', 'ark,t:f', 'ark,scp:f.ark,f.scp', 'ark:| cmd'."""

    def __init__(self, wspecifier):
        import ctypes as C
        from ._lib import lib, last_error
        self._lib, self._err, self._h = lib, last_error, C.c_void_p()
        if lib.xvb_ark_writer_open(C.byref(self._h), wspecifier.encode()) != 0:
            raise KaldiFormatError(last_error())

    def write(self, key, vec):
        v = np.ascontiguousarray(vec, dtype=np.float32).reshape(-1)
        if self._lib.xvb_ark_writer_put_vector(self._h, key.encode("latin1"), v.ctypes.data, v.shape[0]) != 0:
            raise KaldiFormatError(self._err())

    def close(self):
        if self._h:
            h, self._h = self._h, None
            if self._lib.xvb_ark_writer_close(h) != 0:
                raise KaldiFormatError(self._err())

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
