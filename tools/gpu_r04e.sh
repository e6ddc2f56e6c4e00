for S in tma direct; do echo -n "XVB_GEMM_STORE=$S  "; XVB_GEMM_STORE=$S timeout 120 python tools/layer_times.py 2>&1 | tail -1; done
XVB_GEMM_STORE=direct timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "golden or tdnn_gemm" 2>&1 | tail -3
