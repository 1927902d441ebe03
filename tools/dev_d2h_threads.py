"""Host-side ceiling of concurrent device->host copies on one box, without any kernel running:
ONE process drives every visible GPU from its own thread (own stream, own pinned buffer); reports GB/s per GPU alone,
per GPU with all copying at once, and the sum. bench.py's `e2e.pcie` measures the same with one PROCESS per GPU
(torchrun), so the two launch models can be compared on the same box.

    python tools/dev_d2h_threads.py [--mib 256] [--reps 8]      # prints one JSON line
"""
import argparse
import json
import threading

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=256)
    ap.add_argument("--reps", type=int, default=8)
    a = ap.parse_args()
    n = torch.cuda.device_count()
    size = a.mib << 20
    bufs = []
    for d in range(n):
        with torch.cuda.device(d):
            bufs.append((torch.empty(size, dtype=torch.uint8, device=f"cuda:{d}"), torch.empty(size, dtype=torch.uint8).pin_memory(),
                         torch.cuda.Stream(device=d)))

    def copy_loop(d, reps, out, barrier=None):
        dev, host, stream = bufs[d]
        with torch.cuda.device(d), torch.cuda.stream(stream):
            host.copy_(dev, non_blocking=True)                      # warm-up
            stream.synchronize()
            if barrier is not None:
                barrier.wait()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                host.copy_(dev, non_blocking=True)
            e1.record(stream)
            stream.synchronize()
            out[d] = reps * size / 1e6 / e0.elapsed_time(e1)        # GB/s

    alone = {}
    for d in range(n):
        copy_loop(d, a.reps, alone)
    together = {}
    bar = threading.Barrier(n)
    ts = [threading.Thread(target=copy_loop, args=(d, a.reps, together, bar)) for d in range(n)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    print(json.dumps({"gpus": n, "mib": a.mib, "reps": a.reps, "launch": "one process, one thread per GPU",
                      "d2h_gbs_alone": [round(alone[d], 1) for d in range(n)],
                      "d2h_gbs_concurrent": [round(together[d], 1) for d in range(n)],
                      "d2h_gbs_concurrent_sum": round(sum(together.values()), 1),
                      "d2h_gbs_concurrent_min": round(min(together.values()), 1)}))


if __name__ == "__main__":
    main()
