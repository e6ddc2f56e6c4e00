# -*- coding:utf-8 -*-
"""Waveform-in extractor -- CLI twin of pytorch/pipeline/onestep/extract_embeddings_online.py (:25-135):

    python -m asv_subtools_b200.pipeline.extract_embeddings_online [--nnet-config F | --model-blueprint P
        --model-creation S] [--use-gpu true] [--gpu-id ID] --feat-config feat_conf.yaml [--max-chunk N]
        <model-path> <wav.scp> <vectors-wspecifier>

`feat_conf.yaml` is the file the launchers write next to the model (runEcapaXvector_online.py:378-385;
runtime/test/feat_conf.yaml): `feature_type`, `kaldi_featset`, `mean_var_conf`.  The reference decodes
each file with torchaudio, scales to int16 range (processor.py:429), computes KaldiFeature on the CPU
and extracts one utterance at a time; here PCM16 files are read with the standard library, features for a
whole bucket of files with the same FRAME count come from ONE (ragged) xvb_fbank_compute launch, and the bucket is extracted
in one batched call.  One `FV` vector per key; `RTF:` line at the end like the reference (:131)."""
import argparse
import os
import sys
import time
import traceback
import wave

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from asv_subtools_b200 import kaldi_io  # noqa: E402
from asv_subtools_b200.pipeline.extract_embeddings import Batcher, create_model_from_py, read_nnet_config  # noqa: E402


def read_pcm16(path):
    """-> (float32 samples of channel 0 in int16 range, sample rate)."""
    with wave.open(path, "rb") as w:
        if w.getsampwidth() != 2 or w.getcomptype() != "NONE":
            raise ValueError("{}: only 16-bit PCM wav files are supported".format(path))
        raw = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")
        return raw.reshape(-1, w.getnchannels())[:, 0].astype(np.float32), w.getframerate()


def main(argv=None):
    import yaml
    ap = argparse.ArgumentParser(description="Extract embeddings from a wav.scp (features on the GPU)")
    ap.add_argument("--nnet-config", type=str, default="")
    ap.add_argument("--model-blueprint", type=str, default=None)
    ap.add_argument("--model-creation", type=str, default=None)
    ap.add_argument("--use-gpu", type=str, default="true", choices=["true", "false"])
    ap.add_argument("--gpu-id", type=str, default="")
    ap.add_argument("--feat-config", type=str, required=True)
    ap.add_argument("--max-chunk", type=int, default=10000)
    ap.add_argument("--batch-size", type=int, default=64)
    ap.add_argument("model_path", metavar="model-path")
    ap.add_argument("wav_scp", metavar="wav-scp")
    ap.add_argument("vectors_wspecifier", metavar="vectors-wspecifier")
    print(" ".join(sys.argv))
    args = ap.parse_args(argv)
    try:
        if args.nnet_config != "":
            blueprint, creation = read_nnet_config(args.nnet_config)
        elif args.model_blueprint is not None and args.model_creation is not None:
            blueprint, creation = args.model_blueprint, args.model_creation
        else:
            raise ValueError("Expected nnet_config or (model_blueprint, model_creation) to exist.")
        if args.use_gpu != "true":
            raise RuntimeError("asv_subtools_b200 has no CPU path: run with --use-gpu true on a B200")
        from asv_subtools_b200.frontend import KaldiFeature
        with open(args.feat_config) as f:
            conf = yaml.safe_load(f)
        model = create_model_from_py(blueprint, creation)
        model.load_state_dict(torch.load(args.model_path, map_location="cpu"), strict=False)
        torch.cuda.set_device(int(args.gpu_id.split(",")[0]) if args.gpu_id != "" else 0)
        model.cuda().eval()
        feature, rate = None, None
        # bucket by FRAME count (real files almost never share a sample count); at most ~2 M frames (a few
        # hundred MB of samples) wait for their bucket to fill before everything pending is extracted
        batcher = Batcher(args.batch_size, max_pending_frames=2_000_000, length=lambda wav: feature.num_frames(wav.shape[0]))
        total_dur, t_extract = 0.0, 0.0

        def run(bucket, w):
            nonlocal total_dur, t_extract
            t0 = time.perf_counter()
            feats, fo = feature.compute([wav for _, wav in bucket])
            frames = int(fo[1])
            if frames <= args.max_chunk:
                emb = model.extract_embedding_batch(feats.view(len(bucket), frames, feature.dim)).cpu().numpy()
            else:   # the reference's chunk rule, one utterance at a time
                emb = np.stack([model.extract_embedding(feats[i * frames:(i + 1) * frames].cpu().numpy()).numpy()
                                for i in range(len(bucket))])
            t_extract += time.perf_counter() - t0
            total_dur += len(bucket) * frames * 0.01
            for (key, _), e in zip(bucket, emb):
                kaldi_io.write_vec_flt(w, e, key=key)

        with open(args.wav_scp) as scp, kaldi_io.open_or_fd(args.vectors_wspecifier, "wb") as w:
            for line in scp:
                parts = line.strip().split(None, 1)
                if len(parts) != 2:
                    continue
                key, path = parts
                wav, sr = read_pcm16(path)
                if feature is None:
                    rate = sr
                    featset = dict(conf.get("kaldi_featset", {}))
                    featset["sample_frequency"] = float(sr)        # processor.py:426
                    featset["dither"] = 0.0                        # forced for extraction, runEcapaXvector_online.py:381
                    feature = KaldiFeature(conf.get("feature_type", "fbank"), featset, conf.get("mean_var_conf", {}))
                elif sr != rate:
                    raise ValueError("{} is sampled at {} Hz, the first file at {} Hz".format(key, sr, rate))
                nf = feature.num_frames(wav.shape[0])
                if nf < 1:
                    raise ValueError("{} is shorter than one analysis window".format(key))
                for bucket in batcher.add(key, wav):
                    run(bucket, w)
            for bucket in batcher.flush():
                run(bucket, w)
        print("RTF:{:.7f}".format(t_extract / max(total_dur, 1e-9)))
    except BaseException as err:
        if not isinstance(err, KeyboardInterrupt):
            traceback.print_exc()
        sys.exit(1)


if __name__ == "__main__":
    main()
