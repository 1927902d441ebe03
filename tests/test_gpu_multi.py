"""Frame sharding over the GPUs of one box on REAL devices: two ranks under torchrun, each with its own engine
replica and host pipeline, no collective on the data path, host-side gather — the gathered per-frame results must be
identical to one GPU processing every frame. Skipped on a single-GPU box (the gloo / CPU version of the plumbing is
tests/test_host.py::test_run_sharded_gloo_world2)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import torch
import torch.distributed as dist
from densepose_torchscript_b200 import synth
from densepose_torchscript_b200.config import BUILTIN
from densepose_torchscript_b200.engine import Engine, HostPipeline
from densepose_torchscript_b200.parallel import run_sharded

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
spec = BUILTIN["densepose_rcnn_R_50_FPN_s1x_legacy"]
eng = Engine(spec, synth.make_state_dict(spec, 0), device=torch.device("cuda", local))
frames = [synth.synthetic_image(160, 256, seed=300 + i).round().clamp(0, 255).to(torch.uint8) for i in range(22)]
B = 4

def make_process(engine):
    pipe = HostPipeline(engine, B, 160, 256, src_u8=True, depth=2, extract=True)
    def process(batch):
        pad = batch + [batch[-1]] * (B - len(batch))                  # the last batch of a shard may be short
        assert pipe.submit(torch.stack(pad)) is None
        (res,) = pipe.drain()
        return [{"boxes": r["pred_boxes"].clone(), "scores": r["scores"].clone(),
                 "labels": [d["labels"].clone() for d in r["densepose"]], "rank": rank} for r in res[:len(batch)]]
    return process

got = run_sharded(frames, make_process(eng), batch=B)
if rank == 0:
    assert {g["rank"] for g in got} == {0, 1}
    single = make_process(eng)
    want = []
    for i in range(0, len(frames), B):
        want += single(frames[i:i + B])
    assert len(got) == len(want) == 22
    for g, w in zip(got, want):
        assert torch.equal(g["boxes"], w["boxes"]) and torch.equal(g["scores"], w["scores"])
        assert len(g["labels"]) == len(w["labels"]) and all(torch.equal(a, b) for a, b in zip(g["labels"], w["labels"]))
    print("SHARDED_GPU_OK", sum(len(g["labels"]) for g in got))
else:
    assert got is None
dist.barrier()
dist.destroy_process_group()
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on the box (gpurun --gpus 2)")
def test_frames_sharded_over_two_gpus_equal_one_gpu(tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29741", str(script), root]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    assert "SHARDED_GPU_OK" in r.stdout
