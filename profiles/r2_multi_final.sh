#!/bin/bash
# final multi-GPU record on one 8-GPU box: collected multi-GPU tests, then the bench at N = 1, 2, 4, 8 back to back
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2_multi_tests.log 2>&1; echo "multi tests rc=$?"; tail -3 gpurun_out/r2_multi_tests.log
python bench.py --gpus 1 --steps 20 --warmup 3 --no-eval --no-cpu-baseline > gpurun_out/r2m_bench_1gpu.json 2> gpurun_out/r2m_1.err; echo "N=1 rc=$?"
for N in 2 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 20 --warmup 3 --no-eval > gpurun_out/r2m_bench_${N}gpu.json 2> gpurun_out/r2m_$N.err; echo "N=$N rc=$?"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2m_bench_8gpu.json 2> gpurun_out/r2m_8.err; echo "N=8 rc=$?"
python - <<'PY'
import json
for n in (1,2,4,8):
    try:
        d=json.loads(open(f"gpurun_out/r2m_bench_{n}gpu.json").read().strip().splitlines()[-1])
        g=d.get("global_batch_256") or {}
        e=d.get("eval") or {}
        print(n, round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "g256", g.get("value"), g.get("ms_per_step"), "eval", e.get("value"))
    except Exception as ex: print(n, "failed", ex)
PY
