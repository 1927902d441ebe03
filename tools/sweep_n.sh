#!/bin/bash
# BASELINE.json configs 2-5 on N GPUs of one box (weak scaling: every rank runs its own batch, no collective on the data
# path): one bench.py JSON line per run, appended to <out>. Usage (under gpurun --gpus N):
#   tools/sweep_n.sh <N> <out.jsonl> [quick]
# N = 1 runs bench.py directly, N > 1 under torchrun exactly like the driver does.
n=${1:-1}
out=${2:-gpurun_out/sweep_n${n}.jsonl}
quick=${3:-}
: > "$out"
port=29700
run() {
  port=$((port + 1))
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline --no-latency "$@" 2>>"${out%.jsonl}.err" | tail -1 >> "$out"
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus "$n" --steps 8 --warmup 3 --no-cpu-baseline --no-latency "$@" 2>>"${out%.jsonl}.err" | grep '^{' | tail -1 >> "$out"
  fi
}
# configs[1] / [4]: R50-s1x, batch 8 with the host pipelines (e2e, its variants, the 10-detection operating point) ...
run --batch 8
# configs[2]: R101 DeepLab head, batch 32 over 8 GPUs = 4 per GPU
run --config densepose_rcnn_R_101_FPN_DL_s1x --batch 4 --realistic-dets 0
# configs[3]: R101-s1x on 1080p frames with ~100 boxes per frame
run --config densepose_rcnn_R_101_FPN_s1x --batch 4 --height 1080 --width 1920 --realistic-dets 0
if [ -z "$quick" ]; then
  # configs[4]: the batch sweep (device-resident number only)
  batches="1 2 4 16 32 64"
  [ "$n" != 1 ] && batches="1 32"          # N GPUs cost N x the box time: the ends of the sweep are enough
  for b in $batches; do run --batch $b --no-e2e --realistic-dets 0; done
  [ "$n" = 1 ] && run --config densepose_rcnn_R_50_FPN_s1x_legacy --batch 8 --no-e2e --realistic-dets 0
fi
python - "$out" <<'P'
import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l)
    e = d.get("e2e", {})
    r = d.get("realistic_dets", {})
    print(d["config"]["workload"][:60], "| N", d["n_gpus"], "B", d["config"]["batch_per_gpu"], "|", round(d["value"], 1), "img/s |",
          round(d["ms_per_step"], 2), "ms |", round(d["roofline"]["step_tflops"]), "TF/s/GPU | e2e", round(e.get("value", 0), 1),
          "| x", round(e.get("variants", {}).get("extracted", {}).get("value", 0), 1),
          "| D10", round(r.get("value", 0), 1), "e2e", round(r.get("e2e", {}).get("value", 0), 1))
P
