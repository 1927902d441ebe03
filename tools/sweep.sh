#!/bin/bash
# BASELINE.json configs 3-5 on ONE GPU (device-resident throughput, CUDA events): batch sweep of R50-s1x, the R101
# DeepLab head at its per-GPU batch, R101-s1x on 1080p frames, and the legacy head. One JSON line per run.
out=${1:-gpurun_out/sweep.jsonl}
: > "$out"
run() { python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-latency --no-e2e "$@" 2>/dev/null | tail -1 >> "$out"; }
for b in 1 2 4 8 16 32; do run --batch $b; done
run --config densepose_rcnn_R_101_FPN_DL_s1x --batch 4
run --config densepose_rcnn_R_101_FPN_s1x --batch 4 --height 1080 --width 1920
run --config densepose_rcnn_R_50_FPN_s1x_legacy --batch 8
python - "$out" <<'P'
import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l)
    print(d["config"]["workload"][:70], "| B", d["config"]["batch_per_gpu"], "|", round(d["value"], 1), "img/s |", round(d["ms_per_step"], 2), "ms |",
          round(d["roofline"]["step_tflops"]), "TF/s step")
P
