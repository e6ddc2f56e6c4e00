timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_ecapa.py -m gpu -q -x --durations=5 -k "lde or snowdar or small_affine or ecapa" 2>&1 | tail -12
