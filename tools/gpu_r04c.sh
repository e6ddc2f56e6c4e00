for D in 1 3; do echo "== XVB_RES2_DEBUG=$D"; XVB_RES2_DEBUG=$D timeout 120 python tools/bench_ecapa.py 10 --profile 2>&1 | grep -E "res2net|sum|frames_per_s|Error|error" | cut -c1-200; done
