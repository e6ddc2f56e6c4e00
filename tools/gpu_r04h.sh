TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 tools/peer_table_check.py 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tail -12
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "replicated or shard" 2>&1 | tail -3
for G in p2p nccl; do
  XVB_BENCH_GATHER=$G XVB_BENCH_UTTS=32768 XVB_BENCH_ECAPA_UTTS=4096 timeout 600 $TR --master-port 29522 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r04h_bench_2gpu_$G.json 2> gpurun_out/r04h_bench_2gpu_$G.err; echo "gather=$G rc=$?"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r04h_bench_2gpu_$G.err | tail -4
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r04h_bench_2gpu_$G.json").read().strip().splitlines()[-1])
    print("$G value %.4e ms/step %.2f" % (d["value"], d["ms_per_step"]), d["exchange"]["kind"], d["exchange"]["p2p_equals_nccl"], {k: v for k, v in d["phases_ms"].items() if k != "note"}, "c4 eq", d["config4"].get("eer_equals_single_gpu"))
except Exception as e:
    print("parse failed", e)
PY
done
