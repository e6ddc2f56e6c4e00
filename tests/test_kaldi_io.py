"""asv_subtools_b200.kaldi_io against byte streams written / decoded by the reference's own
kaldi_io (tests/golden/kaldi_ark.npz).  CPU only."""
import io
import os

import numpy as np
import pytest

from asv_subtools_b200 import kaldi_io


def test_matrices_read_and_write_are_byte_identical(golden):
    g = golden("kaldi_ark")
    blob = g["ark_mats_bytes"].tobytes()
    got = dict(kaldi_io.read_mat_ark(io.BytesIO(blob)))
    assert list(got) == ["utt-a", "utt_b"]
    assert got["utt-a"].dtype == np.float32 and np.array_equal(got["utt-a"], g["mat32"])
    assert got["utt_b"].dtype == np.float64 and np.array_equal(got["utt_b"], g["mat64"])
    out = io.BytesIO()
    kaldi_io.write_mat(out, g["mat32"], key="utt-a")
    kaldi_io.write_mat(out, g["mat64"], key="utt_b")
    assert out.getvalue() == blob


def test_vectors_read_and_write_are_byte_identical(golden):
    g = golden("kaldi_ark")
    blob = g["ark_vecs_bytes"].tobytes()
    got = dict(kaldi_io.read_vec_flt_ark(io.BytesIO(blob)))
    assert np.array_equal(got["spk1"], g["vec32"]) and np.array_equal(got["spk2"], g["vec64"])
    out = io.BytesIO()
    kaldi_io.write_vec_flt(out, g["vec32"], key="spk1")
    kaldi_io.write_vec_flt(out, g["vec64"], key="spk2")
    assert out.getvalue() == blob


def test_compressed_matrix_decodes_like_the_reference(golden):
    g = golden("kaldi_ark")
    fd = io.BytesIO(g["ark_cm_bytes"].tobytes())
    assert kaldi_io.read_key(fd) == "cmutt"
    m = kaldi_io.read_mat(fd)
    assert m.dtype == np.float32 and m.shape == g["ark_cm_decoded"].shape
    assert np.allclose(m, g["ark_cm_decoded"], rtol=1e-6, atol=1e-6)
    assert kaldi_io.read_key(fd) is None


def test_ark_scp_roundtrip_with_offsets_and_prefixes(tmp_path):
    rng = np.random.RandomState(0)
    items = [("k{}".format(i), rng.standard_normal(5 + i).astype(np.float32)) for i in range(4)]
    ark, scp = str(tmp_path / "v.ark"), str(tmp_path / "v.scp")
    kaldi_io.write_vec_ark_scp(ark, scp, items)
    a = dict(kaldi_io.read_vec_flt_ark("ark:" + ark))
    s = dict(kaldi_io.read_vectors("scp:" + scp))
    for k, v in items:
        assert np.array_equal(a[k], v) and np.array_equal(s[k], v)
    feats = str(tmp_path / "f.ark")
    with open(feats, "wb") as f:
        for k, v in items:
            kaldi_io.write_mat(f, np.tile(v, (3, 1)), key=k)
    m = dict(kaldi_io.read_mat_ark(feats))
    assert m["k2"].shape == (3, 7)
    # input pipe rspecifier, like the reference's "ark:copy-feats ... |"
    piped = dict(kaldi_io.read_mat_ark("ark:cat {} |".format(feats)))
    assert np.array_equal(piped["k3"], m["k3"])


def test_ascii_forms():
    m = kaldi_io.read_mat(io.BytesIO(b" [\n 1 2 3\n 4 5 6 ]\n"))
    assert np.array_equal(m, np.array([[1, 2, 3], [4, 5, 6]], dtype=np.float32))
    v = kaldi_io.read_vec_flt(io.BytesIO(b" [ 1.5 2.5 ]\n"))
    assert np.array_equal(v, np.array([1.5, 2.5]))


def test_errors():
    with pytest.raises(kaldi_io.KaldiFormatError):
        kaldi_io.read_mat(io.BytesIO(b"\0BXM \4\1\0\0\0\4\1\0\0\0"))
    with pytest.raises(kaldi_io.KaldiFormatError):
        kaldi_io.read_mat(io.BytesIO(b"\0BFM \4\2\0\0\0\4\2\0\0\0\0\0"))  # truncated payload
    with pytest.raises(TypeError):
        kaldi_io.write_vec_flt(io.BytesIO(), np.arange(3))


# ------------------------------------------------------------------ native C reader / writer (csrc/ark_io.cpp)
def _compressed_ark(key, rows, cols, seed):
    """A hand-built 'CM ' entry: header + per-column percentiles + column-major bytes."""
    import struct
    rng = np.random.RandomState(seed)
    gmin, grange = np.float32(-3.25), np.float32(11.5)
    perc = np.sort(rng.randint(0, 65536, size=(cols, 4)).astype("<u2"), axis=1)
    data = rng.randint(0, 256, size=(cols, rows)).astype(np.uint8)
    data[0, :4] = [0, 64, 192, 255]
    data[-1, :3] = [65, 193, 128]
    return (key + " ").encode() + b"\0BCM " + struct.pack("<ffii", gmin, grange, rows, cols) + perc.tobytes() + data.tobytes()


def test_native_reader_matches_python_reader_bit_for_bit(golden, tmp_path):
    g = golden("kaldi_ark")
    ark = tmp_path / "f.ark"
    blob = g["ark_mats_bytes"].tobytes() + g["ark_cm_bytes"].tobytes() + _compressed_ark("cm2", 37, 5, 3) + \
        b"txt  [\n  1.5 -2 3e-3\n  4 5 6 ]\n" + b"empty  [ ]\n"
    ark.write_bytes(blob)
    py = list(kaldi_io.read_mat_ark(str(ark)))
    for spec in (str(ark), "ark:" + str(ark), "ark:cat {} |".format(ark)):
        nat = list(kaldi_io.read_mat_ark_native(spec))
        assert [k for k, _ in nat] == [k for k, _ in py] == ["utt-a", "utt_b", "cmutt", "cm2", "txt", "empty"]
        for (k, a), (_, b) in zip(nat, py):
            assert a.dtype == b.dtype and a.shape == b.shape, k      # 'DM ' stays float64 like read_mat
            assert np.array_equal(a.astype(np.float32), b.astype(np.float32)), k   # incl. the CM decode: same fp32 steps
    assert np.array_equal(dict(py)["cmutt"], g["ark_cm_decoded"]) or np.allclose(dict(py)["cmutt"], g["ark_cm_decoded"], rtol=1e-6)
    # scp with byte offsets and a pipe entry
    offs, pos = {}, 0
    fd = open(ark, "rb")
    while True:
        k = kaldi_io.read_key(fd)
        if k is None:
            break
        offs[k] = fd.tell()
        kaldi_io.read_mat(fd)
    scp = tmp_path / "f.scp"
    scp.write_text("utt_b {a}:{o1}\ncm2 {a}:{o2}\n\nutt-a tail -c +{o3} {a} | \n".format(a=ark, o1=offs["utt_b"], o2=offs["cm2"], o3=offs["utt-a"] + 1))
    nat = list(kaldi_io.read_mat_ark_native("scp:" + str(scp)))
    assert [k for k, _ in nat] == ["utt_b", "cm2", "utt-a"]
    assert np.array_equal(nat[0][1], g["mat64"].astype(np.float32)) and np.array_equal(nat[1][1], dict(py)["cm2"])


def test_native_reader_rejects_malformed_streams(tmp_path):
    bad = tmp_path / "bad.ark"
    for blob in (b"k \0BXM \4\1\0\0\0\4\1\0\0\0\0\0\0\0", b"k \0BFM \4\5\0\0\0\4\5\0\0\0\1\2", b"k \0BCM2" + b"\0" * 10,
                 b"k \0BCM7 " + b"\0" * 16, b"k zz"):
        bad.write_bytes(blob)
        with pytest.raises(kaldi_io.KaldiFormatError):
            list(kaldi_io.read_mat_ark_native(str(bad)))
    with pytest.raises(kaldi_io.KaldiFormatError):
        list(kaldi_io.read_mat_ark_native(str(tmp_path / "missing.ark")))


def test_native_writer_is_byte_identical_and_scp_offsets_resolve(golden, tmp_path):
    g = golden("kaldi_ark")
    rng = np.random.RandomState(1)
    items = [("spk1", g["vec32"]), ("k2", rng.standard_normal(512).astype(np.float32)), ("k3", np.zeros(0, np.float32))]
    ark, scp = tmp_path / "v.ark", tmp_path / "v.scp"
    with kaldi_io.NativeVectorWriter("ark,scp:{},{}".format(ark, scp)) as w:
        for k, v in items:
            w.write(k, v)
    ref = io.BytesIO()
    for k, v in items:
        kaldi_io.write_vec_flt(ref, v, key=k)
    assert ark.read_bytes() == ref.getvalue()
    assert ark.read_bytes().startswith(g["ark_vecs_bytes"].tobytes()[:4 + 1 + 2 + 3 + 5 + 44])   # the reference's own bytes for spk1
    back = dict(kaldi_io.read_vectors("scp:" + str(scp)))
    for k, v in items[:2]:
        assert np.array_equal(back[k], v)
    # text mode and an output pipe
    with kaldi_io.NativeVectorWriter("ark,t:" + str(tmp_path / "t.ark")) as w:
        w.write("a", np.array([1.5, -2.0, 1e-7], np.float32))
    assert (tmp_path / "t.ark").read_text() == "a  [ 1.5 -2 1.00000001e-07 ]\n"
    assert np.array_equal(dict(kaldi_io.read_vec_flt_ark(str(tmp_path / "t.ark")))["a"].astype(np.float32),
                          np.array([1.5, -2.0, 1e-7], np.float32))   # %.9g round-trips fp32
    with kaldi_io.NativeVectorWriter("ark:| cat > {}".format(tmp_path / "p.ark")) as w:
        w.write("k2", items[1][1])
    assert np.array_equal(dict(kaldi_io.read_vec_flt_ark(str(tmp_path / "p.ark")))["k2"], items[1][1])
    with pytest.raises(kaldi_io.KaldiFormatError):
        kaldi_io.NativeVectorWriter("scp:" + str(tmp_path / "x.scp"))
    with kaldi_io.NativeVectorWriter("ark:" + str(tmp_path / "e.ark")) as w:
        with pytest.raises(kaldi_io.KaldiFormatError):
            w.write("two words", items[0][1])


def test_native_reader_decodes_kaldi_cm2_cm3(tmp_path):
    """The header-only compressed formats Kaldi also writes (the reference's reader rejects them; the native one,
    which replaces the `copy-feats ark:- |` pipe in front of the extractor, reads them)."""
    import struct
    rng = np.random.RandomState(4)
    gmin, grange = np.float32(-7.5), np.float32(21.25)
    q16 = rng.randint(0, 65536, size=(9, 5)).astype("<u2")
    q8 = rng.randint(0, 256, size=(4, 6)).astype(np.uint8)
    blob = b"a \0BCM2 " + struct.pack("<ffii", gmin, grange, 9, 5) + q16.tobytes() + \
           b"b \0BCM3 " + struct.pack("<ffii", gmin, grange, 4, 6) + q8.tobytes()
    p = tmp_path / "c.ark"
    p.write_bytes(blob)
    got = dict(kaldi_io.read_mat_ark_native(str(p)))
    assert np.allclose(got["a"], gmin + grange * q16.astype(np.float32) / np.float32(65535), rtol=1e-6, atol=1e-6)
    assert np.allclose(got["b"], gmin + grange * q8.astype(np.float32) / np.float32(255), rtol=1e-6, atol=1e-6)
    with pytest.raises(kaldi_io.KaldiFormatError):               # the Python mirror keeps the reference's behaviour
        list(kaldi_io.read_mat_ark(str(p)))
