#!/bin/bash
# The end-of-round scaling runs on one 8 x B200 box: bench.py at N = 8, 4, 2, 1 (default mode), NV12 ingest at N = 8,
# and BASELINE configs[4] (64 4K streams over 8 GPUs).  usage: gpurun --gpus 8 -- bash tools/scale_run.sh
mkdir -p gpurun_out
run() { # n, tag, extra args...
  n=$1; tag=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      bench.py --gpus $n "$@" > gpurun_out/r02_bench_$tag.json 2> gpurun_out/r02_bench_$tag.err || tail -5 gpurun_out/r02_bench_$tag.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_bench_$tag.json"))
    print("$tag", "value %.0f" % d["value"], d["unit"], "| e2e %.0f" % d["e2e"]["value"], "| ms/step %.2f" % d["ms_per_step"],
          "| shares", d["e2e"].get("frames_per_rank"), "|", d["config"].get("placement", ""))
except Exception as e:
    print("$tag failed:", e)
PY
}
run 8 n8 --steps 10 --warmup 3 --no-extras --no-parity --no-cpu
run 4 n4 --steps 10 --warmup 3 --no-extras --no-parity --no-cpu
run 2 n2 --steps 10 --warmup 3 --no-extras --no-parity --no-cpu
python bench.py --steps 10 --warmup 3 --no-extras --no-parity --no-cpu > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1.json')); print('n1 value %.0f | e2e %.0f' % (d['value'], d['e2e']['value']))"
run 8 n8_nv12 --steps 10 --warmup 3 --no-extras --no-parity --no-cpu --ingest nv12
run 8 n8_streams4k --config streams4k --steps 20 --warmup 3
