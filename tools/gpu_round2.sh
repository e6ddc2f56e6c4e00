#!/bin/bash
# One GPU-box visit of round 2.  Usage (under gpurun): bash tools/gpu_round2.sh <tag> [tests] [bench] [small] [ncu]
TAG=${1:-r03}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt
for WHAT in "$@"; do
case $WHAT in
tests)
  timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest_gpu.log
  timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log ;;
small)
  XVB_BENCH_UTTS=8192 XVB_BENCH_ECAPA_UTTS=2048 timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_small.json 2> gpurun_out/${TAG}_bench_small.err; echo "small bench rc=$?"; tail -c 3000 gpurun_out/${TAG}_bench_small.json; tail -5 gpurun_out/${TAG}_bench_small.err ;;
bench)
  timeout 900 python bench.py --steps ${STEPS:-10} --warmup ${WARMUP:-3} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/${TAG}_bench.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("value %.4e e2e %.4e ms/step %.2f frac %.3f exec %.3f | burst %.4e frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["executed_frac"], d["burst"]["value"], d["burst"]["roofline"]["frac"]))
    print("kernel_ms", d["kernel_ms"]); print("burst kernel_ms", d["burst"]["kernel_ms"]); print("clocks", d["clocks"]); print("e2e", {k: v for k, v in d["e2e"].items() if k not in ("api",)})
    print("c4", d["config4"]); print("c3", d["config3_ecapa"]); print("c5", d["config5"]); print("cpu", d.get("cpu_baseline")); print("c1", d.get("c1_single_utterance")); print("pool", d["roofline_stats_pool"]["frac"])
except Exception as e:
    print("bench parse failed", e)
PY
  ;;
ref)
  timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/${TAG}_bench_ref.json ;;
ncu)
  XVB_BENCH_UTTS=2048 XVB_BENCH_ECAPA_UTTS=512 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu launches rc=$?" ;;
ncufull)
  XVB_BENCH_UTTS=2048 XVB_BENCH_ECAPA_UTTS=512 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tdnn_gemm -s 60 -c 6 -o gpurun_out/${TAG}_gemm python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?" ;;
esac
done
ls -la gpurun_out | tail -12
