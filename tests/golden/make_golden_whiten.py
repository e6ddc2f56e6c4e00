#!/usr/bin/env python
"""Golden ZCA whitening matrix from the REFERENCE's own score/whiten/train_ZCA_Whitening.py, run as the shell
function `trainwhiten` runs it (score/process.sh:235-248): text ark in, Kaldi text matrix out.  Build container
only:   python tests/golden/make_golden_whiten.py  ->  tests/golden/whiten.npz"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import scoring as osc  # noqa: E402


def main():
    emb, _ = osc.synthetic_speakers(30, 8, 24, 41, noise=0.8)
    emb = (emb * np.linspace(0.3, 2.0, 24)[None, :] + 0.4).astype(np.float32)      # anisotropic, non-zero mean
    with tempfile.TemporaryDirectory() as d:
        txt, mat = os.path.join(d, "train.ark.txt"), os.path.join(d, "whiten.mat")
        with open(txt, "w") as f:
            for i, v in enumerate(emb):
                f.write("u%03d  [ %s ]\n" % (i, " ".join(repr(float(x)) for x in v)))
        subprocess.run([sys.executable, "/root/reference/score/whiten/train_ZCA_Whitening.py", "--ark-format=true", txt, mat],
                       check=True, stdout=subprocess.DEVNULL)
        rows = [l.split() for l in open(mat).read().replace("[", " ").replace("]", " ").splitlines() if l.split()]
        w = np.array(rows, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "whiten.npz"), emb=emb, zca=w)
    print("whiten.npz", w.shape, w[0, :3])


if __name__ == "__main__":
    main()
