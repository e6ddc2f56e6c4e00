bash tools/gpu_round3.sh r04b tests
for D in 0 1 2 3; do echo "== XVB_RES2_DEBUG=$D"; XVB_RES2_DEBUG=$D timeout 300 python tools/bench_ecapa.py 10 --profile 2>&1 | grep -E "res2net|sum|frames_per_s" | cut -c1-200; done
