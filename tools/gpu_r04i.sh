timeout 900 python -m pytest tests/test_gpu_ecapa.py tests/test_gpu_deploy.py -m gpu -q -x --durations=5 2>&1 | tail -12
for S in 1 0; do echo "== XVB_ECAPA_SMALL=$S"; XVB_ECAPA_SMALL=$S timeout 200 python tools/bench_ecapa.py 10 --profile 2>&1 | grep -E "rows|K=1x1024 N=128|K=1x128 N=1024|K=1x3072 N=1[29]|sum|frames_per_s" | cut -c1-160; done
for S in 1 0; do XVB_ECAPA_SMALL=$S XVB_BENCH_UTTS=8192 XVB_BENCH_ECAPA_UTTS=16384 timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r04i_bench_small$S.json 2> gpurun_out/r04i_bench_small$S.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r04i_bench_small$S.json").read().strip().splitlines()[-1])
print("small=$S ecapa %.4e e2e %.4e" % (d["config3_ecapa"]["value"], d["config3_ecapa"]["e2e"]["value"]))
PY
done
