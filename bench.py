#!/usr/bin/env python
"""bench.py -- frames/sec of x-vector extraction (80-d fbank, 200-frame chunks, batch 256 per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path (split -> tdnn1..5 with the statistics pooling fused into
tdnn5's epilogue -> Chan merge -> tdnn6.affine) over one
batch of 256 x 200 x 80 synthetic frames per GPU (BASELINE.json configs[1]).  One process per GPU;
under torchrun the ranks only share a barrier and a MAX-reduce of the device time (utterances shard
with no data-path collective -> weak scaling).  Rank 0 prints ONE JSON line.

  value     : whole-job frames/s with the inputs resident in HBM (CUDA events, max over ranks)
  e2e       : same through the C-ABI host-buffer call (pinned host feats -> H2D -> extract -> D2H)
  roofline  : the tcgen05 TDNN GEMM -- algorithmic FLOPs (5 630 976 FLOP/frame) / summed
              per-launch CUDA-event durations, against the measured bf16 peak (MEASURED_PEAKS.json);
              the kernel executes 3 bf16 MMAs per algorithmic MAC (bf16x3 split), reported too
  cpu_baseline : the oracle port of the reference's CPU PyTorch path on this box's host cores

`--impl reference` times that CPU port alone (the reference is Python/torch and cannot travel to
the GPU box; the oracle restates it op for op, pinned by tests/golden).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B, T, F, D = 256, 200, 80, 512
FLOP_PER_FRAME = 5630976          # 2*(2 807 808 MAC/frame) + 2*1 536 000/200
GEMM_FLOP_PER_STEP = FLOP_PER_FRAME * B * T
POOL_BYTES_PER_STEP = 1212000 * B  # 4*(C*T + 2C) B/utt, C=1500, T=200
NUM_INPUT_BATCHES = 8             # rotate 8 x 16.4 MB inputs; activations per step ~1.1 GB >> 126 MB L2
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures
NCU_GEMM_TRAFFIC_BYTES_PER_LAUNCH = int((62.9 + 169.6 + 171.9 + 158.7 + 155.9 + 9.3) * 1e6 / 6)  # profiles/r02h_gemm_ncu_summary.txt
# sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active and gpu__time_duration of the same capture
NCU_TENSOR_PIPE = {"tdnn1": (46.5, 79.7), "tdnn2": (86.6, 177.2), "tdnn3": (86.4, 177.9), "tdnn4": (59.5, 81.4),
                   "tdnn5+pool": (83.1, 174.4), "tdnn6": (12.6, 14.9)}
NCU_POOL_TRAFFIC_BYTES_PER_LAUNCH = int((307.2 + 5.8) * 1e6)   # profiles/r01x_pool_ncu_summary.txt
METRIC = "frames/sec x-vector extraction (80-d fbank)"
WORKLOAD = ("x-vector TDNN (pytorch/model/xvector.py), 80-d fbank, 200-frame chunks, batch 256 per GPU, "
            "extracted_embedding=far (BASELINE configs[1])")
UNIT = "frames/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"],
                "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clock/throttle sampling during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                mhz = float(parts[0])
                smax = float(parts[1])
            except ValueError:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(mhz)
                for n, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        if not sm:  # region shorter than the sampling period: take whatever we have
            sm = [float(l.split(",")[0]) for _, l in self.lines if l and l.split(",")[0].strip().replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_checkpoint():
    from oracle import nnet as onn  # synthetic seeded weights of the BASELINE architecture (no datasets here)
    return onn.make_state_dict(onn.xvector_spec(F), 102)


# ------------------------------------------------------------------------------------------ CPU arm
def best_cpu_threads(fn, candidates=None):
    """The host arm gets the thread count that serves it best: more threads than the container may
    actually schedule (cgroup quota) makes ATen's small convolutions slower, not faster."""
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (candidates or (1, 4, 8, 16, 32, 64, ncpu)) if c <= ncpu})
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        fn()
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def cpu_port_frames_per_s(sd, budget_s, sample_utts, threads):
    """Oracle port of the reference's CPU PyTorch path on a bounded sample of the same workload:
    (i) batched forward (most favourable to the reference), (ii) the reference's literal
    one-utterance-per-call extract_embedding loop."""
    from oracle import nnet as onn
    feats = onn.synthetic_feats(sample_utts, T, F, 1024)
    x = torch.from_numpy(feats).transpose(1, 2).contiguous()
    with torch.no_grad():
        threads = best_cpu_threads(lambda: onn.xvector_forward(sd, x[:16], "far"))
        onn.xvector_forward(sd, x[:8], "far")  # warm-up
        n, t0 = 0, time.perf_counter()
        while True:
            onn.xvector_forward(sd, x, "far")
            n += 1
            if time.perf_counter() - t0 > budget_s * 0.6:
                break
        batched = n * sample_utts * T / (time.perf_counter() - t0)
        best_cpu_threads(lambda: onn.extract_embedding(lambda z: onn.xvector_forward(sd, z, "far"), feats[0]))
        m, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < budget_s * 0.4:
            onn.extract_embedding(lambda z: onn.xvector_forward(sd, z, "far"), feats[m % sample_utts])
            m += 1
        per_utt = m * T / (time.perf_counter() - t0)
    return batched, per_utt, threads


def run_reference(args, rank, world):
    if rank != 0:
        return
    sd = make_checkpoint()
    from oracle import nnet as onn
    sample_utts = 64
    x = torch.from_numpy(onn.synthetic_feats(sample_utts, T, F, 1024)).transpose(1, 2).contiguous()
    with torch.no_grad():
        threads = best_cpu_threads(lambda: onn.xvector_forward(sd, x[:16], "far"))
        for _ in range(max(1, min(args.warmup, 3))):
            onn.xvector_forward(sd, x, "far")
        t0 = time.perf_counter()
        for _ in range(args.steps):
            onn.xvector_forward(sd, x, "far")
        dt = time.perf_counter() - t0
    value = args.steps * sample_utts * T / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "frames_per_utt": T, "feat_dim": F,
                   "parallelism": "host cores of rank 0 ({} threads)".format(threads),
                   "sample": "each step is a bounded sample of the workload: {} of its {} utterances x {} frames, "
                             "batched forward (the form most favourable to the reference)".format(sample_utts, B, T),
                   "weights": "seeded synthetic checkpoint of the reference architecture"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "{} steps x {} utts x {} frames, oracle port (torch CPU ops incl. masked taps) of "
                                   "the reference forward, batched".format(args.steps, sample_utts, T)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def c1_latency(dev):
    """BASELINE configs[0] on the GPU path: one (200, 23) MFCC utterance through the plugin call
    `model.extract_embedding(ndarray) -> CPU tensor` (H2D, 8 kernels, D2H + sync per call), and the same
    utterance through the oracle port on the host cores."""
    from asv_subtools_b200.model.xvector import Xvector
    from oracle import nnet as onn
    sd = onn.make_state_dict(onn.xvector_spec(23), 101)
    m = Xvector(23, 10, training=False, extracted_embedding="far")
    m.load_state_dict(sd, strict=True)
    m.to(dev).eval()
    feats = onn.synthetic_feats(1, 200, 23, 5)[0]
    for _ in range(10):
        m.extract_embedding(feats)
    ts = []
    for _ in range(100):
        t0 = time.perf_counter()
        m.extract_embedding(feats)
        ts.append(time.perf_counter() - t0)
    gpu_ms = statistics.median(ts) * 1e3
    cs = []
    nt = best_cpu_threads(lambda: onn.extract_embedding(lambda z: onn.xvector_forward(sd, z, "far"), feats))
    for _ in range(20):
        t0 = time.perf_counter()
        onn.extract_embedding(lambda z: onn.xvector_forward(sd, z, "far"), feats)
        cs.append(time.perf_counter() - t0)
    return {"workload": "Xvector(23) one 200-frame utterance, extract_embedding(ndarray)->CPU tensor",
            "gpu_ms_per_utt": gpu_ms, "cpu_port_ms_per_utt": statistics.median(cs) * 1e3, "cpu_threads": nt}


# ------------------------------------------------------------------------------------------ GPU arm
def run_native(args, rank, world, local_rank):
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from asv_subtools_b200.model.xvector import Xvector
    sd = make_checkpoint()
    model = Xvector(F, 10, training=False, extracted_embedding="far")
    model.load_state_dict(sd, strict=True)
    model.cuda().eval()
    ex = model.extractor()

    gen = torch.Generator(device=dev)
    gen.manual_seed(1024 + rank)  # the reference's own seed (runXvector.py:176)
    batches = [torch.randn(B, T, F, device=dev, generator=gen) for _ in range(NUM_INPUT_BATCHES)]
    host = [torch.empty(B, T, F, dtype=torch.float32).pin_memory() for _ in range(2)]
    for h, d in zip(host, batches):
        h.copy_(d.cpu())
    host_out = torch.empty(B, D, dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- device-resident throughput -----------------------------------------------------------
    for i in range(args.warmup):
        ex.extract(batches[i % NUM_INPUT_BATCHES])
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    e0.record()
    for i in range(args.steps):
        out = ex.extract(batches[i % NUM_INPUT_BATCHES])
    e1.record()
    barrier()
    wall1 = time.time()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ex.last_launches * args.steps
    assert torch.isfinite(out).all()

    # ---- end to end through the host-buffer C-ABI call ----------------------------------------
    host_outs = [host_out, torch.empty(B, D, dtype=torch.float32).pin_memory()]

    def e2e_loop(n):
        # submit(i) queues H2D (copy stream) + stack + D2H; wait(i-1) hands batch i-1's embeddings to the host
        ex.submit_host(host[0].data_ptr(), B, T, host_outs[0].data_ptr(), 0)
        for i in range(1, n):
            ex.submit_host(host[i % 2].data_ptr(), B, T, host_outs[i % 2].data_ptr(), i % 2)
            ex.wait((i - 1) % 2)
        ex.wait((n - 1) % 2)

    e2e_loop(max(2, min(args.warmup, 3)))
    barrier()
    t_e2e0 = time.perf_counter()
    e2e_loop(args.steps)
    torch.cuda.synchronize()
    e2e_host_ms = (time.perf_counter() - t_e2e0) * 1e3   # every step ends in the host buffer: host clock
    barrier()
    e2e_ms = max_over_ranks(e2e_host_ms)
    # what bounds e2e: the host->device link.  Same pinned buffer, copy alone, CUDA events.
    dst = torch.empty_like(batches[0])
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dst.copy_(host[0], non_blocking=True)
    torch.cuda.synchronize()
    ev0.record()
    for i in range(8):
        dst.copy_(host[i % 2], non_blocking=True)
    ev1.record()
    torch.cuda.synchronize()
    h2d_gbs = 8 * host[0].numel() * 4 / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
    assert torch.isfinite(host_outs[(args.steps - 1) % 2]).all()
    clocks = sampler.stop(wall0, time.time()) if sampler else None

    # ---- per-kernel CUDA-event times (roofline) -------------------------------------------------
    ex.set_profiling(True)
    per = []
    for i in range(max(3, min(args.steps, 10))):
        ex.extract(batches[i % NUM_INPUT_BATCHES])
        per.append(ex.kernel_times_ms())
    ex.set_profiling(False)
    per = np.median(np.array(per), axis=0)  # [split, tdnn1..4, tdnn5 (+fused pooling), pool_finalize, tdnn6]
    names = ["split"] + ["tdnn%d" % (i + 1) for i in range(4)] + ["tdnn5+pool_partials", "pool_finalize", "tdnn6.affine"]
    gemm_ms = float(per[1:6].sum() + per[7])

    # ---- the standalone statistics-pooling kernel, timed by itself on the BASELINE tensor ------------
    # (the product path pools inside tdnn5's epilogue; the north star also asks for this kernel's HBM fraction)
    from asv_subtools_b200 import ops as _ops
    pool_in = [torch.randn(B, T, 1500, device=dev) for _ in range(3)]   # 3 x 307 MB >> L2
    for i in range(3):
        _ops.stats_pool(pool_in[i % 3])
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for i in range(12):
        _ops.stats_pool(pool_in[i % 3])
    p1.record()
    torch.cuda.synchronize()
    pool_ms = p0.elapsed_time(p1) / 12
    del pool_in

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    frames = B * T * args.steps * world
    value = frames / (ms * 1e-3)
    achieved = GEMM_FLOP_PER_STEP / (gemm_ms * 1e-3) / 1e12
    pool_gbs = POOL_BYTES_PER_STEP / (pool_ms * 1e-3) / 1e9
    # CPU arm beside the GPU number: on rank 0 at N=1 only (it costs ~16 s of host time)
    cpu_batched, cpu_per_utt, cpu_threads = cpu_port_frames_per_s(sd, 16.0, 64, os.cpu_count() or 1) if world == 1 else (None, None, None)
    c1 = c1_latency(dev) if world == 1 else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (bf16x3 split operands on tcgen05, fp32 accumulate in TMEM)", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "batch_per_gpu": B, "frames_per_utt": T, "feat_dim": F, "parallelism": "utterance-sharded x%d" % world,
                   "l2_policy": "inputs rotate over %d batches; ~1.1 GB of activations per step >> 126 MB L2" % NUM_INPUT_BATCHES,
                   "weights": "seeded synthetic checkpoint of the reference architecture"},
        "e2e": {"value": frames / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": B * T * F * 4, "d2h_bytes_per_step": B * D * 4,
                "api": "xvb_extractor_submit_host/xvb_extractor_wait (pinned host feats in, host embeddings out; "
                       "H2D of batch i+1 overlaps the kernels of batch i; timed on the host clock)",
                "h2d_link_gbs_measured": h2d_gbs,
                "h2d_link_bound": world * h2d_gbs * 1e9 / (F * 4), "h2d_link_bound_note":
                "frames/s the host->device link alone allows at 320 B/frame (fp32 80-d features) per GPU"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "tdnn_gemm_bf16x3_kernel (6 launches/step; tdnn5 pools over time in its epilogue, tdnn6 is split-K + a reduce)",
                     "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / pk["bf16_sustained"], "traffic": NCU_GEMM_TRAFFIC_BYTES_PER_LAUNCH,
                     "traffic_unit": "bytes/launch (dram read+write, mean of the 6 launches, profiles/r02h_gemm_ncu_summary.txt)",
                     "peak_source": pk["src"] + " bf16 sustained (kernel timed inside a long step)",
                     "algorithmic_flop_per_launch_avg": GEMM_FLOP_PER_STEP / 6,
                     "executed_tflops": 3 * achieved, "executed_frac": 3 * achieved / pk["bf16_sustained"],
                     "note": "3 bf16 MMAs per algorithmic MAC (hi*hi + lo*hi + hi*lo) to hold 1e-4 parity",
                     "ncu_tensor_pipe_pct": dict({k: v[0] for k, v in NCU_TENSOR_PIPE.items()},
                                                 time_weighted=sum(a * b for a, b in NCU_TENSOR_PIPE.values()) /
                                                 sum(b for _, b in NCU_TENSOR_PIPE.values()),
                                                 source="profiles/r02h_gemm_ncu_summary.txt (ncu --set full, one step)"),
                     "gemm_ms_per_step": gemm_ms},
        "roofline_stats_pool": {"bound": "hbm", "kernel": "stats_pool_tma_kernel (standalone, (256,200,1500) fp32, "
                                                          "12 back-to-back launches over 3 rotating inputs)",
                                "achieved": pool_gbs,
                                "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": pool_gbs / pk["hbm_gbs"],
                                "traffic": NCU_POOL_TRAFFIC_BYTES_PER_LAUNCH,
                                "traffic_unit": "bytes/launch (profiles/r01x_pool_ncu_summary.txt); algorithmic 310.3 MB",
                                "ms": pool_ms, "peak_source": pk["src"]},
        "kernel_ms": {n: float(v) for n, v in zip(names, per)},
    }
    if world == 1:
        line["cpu_baseline"] = {"value": cpu_batched, "unit": UNIT, "cores": cpu_threads, "host_cpus": os.cpu_count() or 1, "kind": "port",
                                "sample": "64 utts x 200 frames batched forward for ~10 s (most favourable to the "
                                          "reference); the reference's literal batch-1 extract_embedding loop: "
                                          "%.0f frames/s" % cpu_per_utt,
                                "per_utterance_value": cpu_per_utt}
        line["c1_single_utterance"] = c1
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no GPU visible -- the native arm has no CPU fallback")
        run_native(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
