#!/usr/bin/env python
"""Pin the Kaldi-boundary rows of the back end (submean / norm / speaker mean / cosine, and
AS-norm in its ddof = 0 form) with code the REFERENCE TREE holds itself: subtools2/egrecho/score/{utils,score,
asnorm}.py.  Kaldi's binaries (ivector-subtract-global-mean, ivector-normalize-length, ivector-mean,
ivector-compute-dot-products) are neither vendored nor installed; these files are the reference's own restatement
of the same steps and run here.

    python tests/golden/make_golden_egrecho.py          (build container only: needs /root/reference)

The three files are executed unmodified.  What is stubbed is I/O only: `kaldi_native_io` (absent pip package)
and `egrecho.utils.io` / `egrecho.utils.logging` (their readers sit on kaldi_native_io) are replaced by a
dict-backed vector reader/writer and the text-list helpers; every arithmetic line that produces the fixture --
`compute_mean_stats`, `CosineScore.score` (mean subtraction + torch cosine_similarity + "%.5f"),
`spk_vector_mean`, `compute_cohort_stats` (cosine GEMM, np.partition top-n, np.mean / np.std with ddof = 0),
`ScoreNorm.norm` -- is the reference's.
"""
import importlib.util
import logging
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
EG = os.path.join(REF, "subtools2/egrecho")

TABLES = {}          # "scp path" -> {key: vector}: what the stub reader serves


class KaldiVectorReader:
    def __init__(self, scp):
        self.table = TABLES[str(scp)]

    def read(self, key):
        return self.table[key]

    def close(self):
        pass


class KaldiVectorWriter:
    def __init__(self, wdir, name):
        self.path = str(os.path.join(str(wdir), str(name)))
        TABLES[self.path] = {}

    def write(self, key, vec):
        TABLES[self.path][key] = np.asarray(vec)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        with open(self.path, "w") as f:              # an "scp" the text helpers can list keys from
            for k in TABLES[self.path]:
                f.write("{} mem:{}\n".format(k, k))


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def stub_packages():
    for name in ("egrecho", "egrecho.score", "egrecho.utils"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    kni = types.ModuleType("kaldi_native_io")

    class SequentialFloatVectorReader:               # `vector_mean` only; iterates (key, vec)
        def __init__(self, rspecifier):
            self.table = TABLES[rspecifier.split(":", 1)[1]]

        def __enter__(self):
            return iter(self.table.items())

        def __exit__(self, *a):
            pass
    kni.SequentialFloatVectorReader = SequentialFloatVectorReader
    sys.modules["kaldi_native_io"] = kni
    io = types.ModuleType("egrecho.utils.io")

    def read_lists_lazy(path, vector=False):
        with open(path, encoding="utf-8") as f:
            for line in f:
                yield line.strip().split() if vector else line

    def read_key_first_lists_lazy(path, vector=False):
        for parts in read_lists_lazy(path, vector=True):
            yield (parts[0], parts[1:]) if vector else (parts[0], " ".join(parts[1:]))

    def read_key_first_lists(path, vector=False):
        return list(read_key_first_lists_lazy(path, vector))

    io.KaldiVectorReader, io.KaldiVectorWriter = KaldiVectorReader, KaldiVectorWriter
    io.read_lists_lazy, io.read_key_first_lists_lazy, io.read_key_first_lists = (
        read_lists_lazy, read_key_first_lists_lazy, read_key_first_lists)
    io.buf_count_newlines = lambda p: sum(1 for _ in open(p))
    io.get_filename = lambda p: os.path.basename(str(p))
    io.close_cached_kaldi_handles = lambda: None
    sys.modules["egrecho.utils.io"] = io
    lg = types.ModuleType("egrecho.utils.logging")
    lg.get_logger = lambda *a, **k: logging.getLogger("egrecho")
    sys.modules["egrecho.utils.logging"] = lg


def write_scp(path, table):
    TABLES[path] = table
    with open(path, "w") as f:
        for k in table:
            f.write("{} mem:{}\n".format(k, k))


def main():
    stub_packages()
    utils = load("egrecho.score.utils", os.path.join(EG, "score/utils.py"))
    load("egrecho.score.binary_metrics", os.path.join(EG, "score/binary_metrics.py"))
    score = load("egrecho.score.score", os.path.join(EG, "score/score.py"))
    asnorm = load("egrecho.score.asnorm", os.path.join(EG, "score/asnorm.py"))

    rng = np.random.RandomState(731)
    dim, n_spk, per = 64, 40, 6
    spk = rng.standard_normal((n_spk, dim)).astype(np.float32)
    offset = (0.3 * rng.standard_normal(dim)).astype(np.float32)           # a non-zero global mean
    emb = (spk[np.repeat(np.arange(n_spk), per)] + 0.7 * rng.standard_normal((n_spk * per, dim)) + offset).astype(np.float32)
    keys = ["u%04d" % i for i in range(emb.shape[0])]
    c_spk, c_per = 52, 5                                                    # cohort: 52 speakers, means of 5 utterances
    cemb = (rng.standard_normal((c_spk, dim))[np.repeat(np.arange(c_spk), c_per)]
            + 0.7 * rng.standard_normal((c_spk * c_per, dim)) + offset).astype(np.float32)
    ckeys = ["c%04d" % i for i in range(cemb.shape[0])]
    te = rng.randint(0, emb.shape[0], size=700)
    tt = rng.randint(0, emb.shape[0], size=700)
    keep = te != tt
    te, tt = te[keep][:600], tt[keep][:600]
    lab = (te // per == tt // per).astype(np.int64)
    top_n = 20
    out = dict(emb=emb, cohort_utt=cemb, cohort_spk=np.repeat(np.arange(c_spk), c_per).astype(np.int32),
               trial_e=te.astype(np.int32), trial_t=tt.astype(np.int32), label=lab, top_n=np.int64(top_n))

    with tempfile.TemporaryDirectory() as d:
        eval_scp, coh_scp = os.path.join(d, "xvector.scp"), os.path.join(d, "cohort.scp")
        write_scp(eval_scp, dict(zip(keys, emb)))
        write_scp(coh_scp, dict(zip(ckeys, cemb)))
        # global mean (utils.compute_mean_stats via score.vector_mean: np.save of the mean vector)
        mean_path = score.vector_mean(eval_scp, os.path.join(d, "mean.npy"))
        mean = np.load(mean_path)
        out["mean"] = mean
        # speaker means of the cohort (asnorm.spk_vector_mean: compute_mean_stats per spk2utt line)
        spk2utt = os.path.join(d, "spk2utt")
        with open(spk2utt, "w") as f:
            for s in range(c_spk):
                f.write("s%03d %s\n" % (s, " ".join(ckeys[s * c_per:(s + 1) * c_per])))
        spk_scp = asnorm.spk_vector_mean(coh_scp, spk2utt, os.path.join(d, "spk_cohort.scp"))
        spk_tab = TABLES[str(spk_scp)]
        out["cohort_mean"] = np.stack([spk_tab["s%03d" % s] for s in range(c_spk)])
        # cosine scores of mean-subtracted embeddings (score.CosineScore.score)
        trials = os.path.join(d, "set.trials")
        with open(trials, "w") as f:
            for a, b, l in zip(te, tt, lab):
                f.write("%s %s %s\n" % (keys[a], keys[b], "target" if l else "nontarget"))
        scorer = score.CosineScore(eval_scp, submean_vec=mean, cache_size=0)
        (sf,) = scorer.score(trials, storage_dir=d)
        cos = np.array([float(l.split()[2]) for l in open(sf)], dtype=np.float64)
        out["cosine_5dp"] = cos
        scorer0 = score.CosineScore(eval_scp, cache_size=0)                # no submean
        (sf0,) = scorer0.score(trials, storage_dir=os.path.join(d))
        out["cosine_nosub_5dp"] = np.array([float(l.split()[2]) for l in open(sf0)], dtype=np.float64)
        # the file above was overwritten by the second call (same name): rewrite the submean one for asnorm
        (sf,) = scorer.score(trials, storage_dir=d)
        # cohort statistics, full precision (asnorm.compute_cohort_stats), order of asnorm.norm: sorted unique keys
        e_idx, t_idx = np.unique(te), np.unique(tt)
        cvec = out["cohort_mean"] - mean
        em, es = asnorm.compute_cohort_stats(emb[e_idx] - mean, cvec, top_n=top_n)
        tm, ts = asnorm.compute_cohort_stats(emb[t_idx] - mean, cvec, top_n=top_n)
        out.update(stats_e_idx=e_idx.astype(np.int32), stats_t_idx=t_idx.astype(np.int32), e_mean=em, e_std=es,
                   t_mean=tm, t_std=ts)
        # AS-norm end to end (asnorm.ScoreNorm.norm on the 5-decimal score file)
        snm = asnorm.ScoreNorm(eval_scp, str(spk_scp), top_n=top_n, submean_vec=mean, storage_dir=d)
        (nf,) = snm.norm(sf)
        out["asnorm_5dp"] = np.array([float(l.split()[2]) for l in open(nf)], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "egrecho_backend.npz"), **out)
    print("egrecho_backend.npz: cosine", out["cosine_5dp"][:3], "asnorm", out["asnorm_5dp"][:3], "targets", int(lab.sum()))


if __name__ == "__main__":
    main()
