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
