# timing experiments of the GEMM epilogue (library built with -DXVB_TIMING_EXPERIMENTS; results are wrong in these modes)
for D in 0 1 4 16 32 64 48 112 116; do echo -n "XVB_GEMM_DEBUG=$D  "; XVB_GEMM_DEBUG=$D timeout 120 python tools/layer_times.py 2>&1 | tail -1; done
