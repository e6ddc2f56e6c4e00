#!/bin/bash
# Scratch GPU visit: compute-sanitizer on the SIMT kernels added late (fbank, front-end, score normalisation, PLDA helpers).
TAG=${1:-r02b}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fbank.py tests/test_gpu_frontend.py -m gpu -q -x > gpurun_out/${TAG}_memcheck_frontend.log 2>&1; echo "memcheck fbank/frontend rc=$?"; grep -c "Invalid\|out of bounds" gpurun_out/${TAG}_memcheck_frontend.log; tail -4 gpurun_out/${TAG}_memcheck_frontend.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_fbank.py -m gpu -q -x -k "match_reference or batched" > gpurun_out/${TAG}_racecheck_fbank.log 2>&1; echo "racecheck fbank rc=$?"; grep -i "race\|hazard" gpurun_out/${TAG}_racecheck_fbank.log | head -5; tail -3 gpurun_out/${TAG}_racecheck_fbank.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_scoring.py -m gpu -q -x -k "cross_select or normalization" > gpurun_out/${TAG}_memcheck_snorm.log 2>&1; echo "memcheck snorm rc=$?"; tail -3 gpurun_out/${TAG}_memcheck_snorm.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_plda_train.py -m gpu -q -x -k "golden or coral" > gpurun_out/${TAG}_memcheck_plda.log 2>&1; echo "memcheck plda rc=$?"; tail -3 gpurun_out/${TAG}_memcheck_plda.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_trial_histogram.py -m gpu -q -x -k "row_units or plda_terms" > gpurun_out/${TAG}_memcheck_hist.log 2>&1; echo "memcheck hist rc=$?"; tail -3 gpurun_out/${TAG}_memcheck_hist.log
