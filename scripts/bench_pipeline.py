"""Dev tool: D batches in flight.  D independent TrainHotPath instances (own workspaces / outputs / sampler cells), one CUDA
graph each, replayed round-robin on D streams: step i+1's latency-bound proposal chain runs beside step i's HBM-bound
RoIAlign.  Prints ms per step (total time / steps) for D = 1, 2, 3 and stream-priority variants.
usage: python scripts/bench_pipeline.py [steps]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200det
from b200det import fused, workload

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
dev = torch.device("cuda:0")
B, K = 8, 8
w = workload.config2(B=B, K=K)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
cls, reg = [T(c) for c in w["cls"]], [T(r) for r in w["reg"]]
feats = [T(f).contiguous(memory_format=torch.channels_last) for f in w["feats"]]
gt, gl = T(w["gt"]), T(w["gt_label"])
gcount = torch.full((B,), K, dtype=torch.int32, device=dev)
img_hw = torch.tensor([[800.0, 1333.0]] * B, device=dev)


def build(D, prio):
    hps, graphs, streams = [], [], []
    for d in range(D):
        hp = fused.TrainHotPath(B, w["grids"], dev, gt_ld=K, feat_channels=256, layout=1, overlap=True)
        for _ in range(3):
            hp.step(cls, reg, feats, gt, gcount, gl, img_hw)
        torch.cuda.synchronize()
        side = torch.cuda.Stream(priority=prio[d % len(prio)])
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            hp.step(cls, reg, feats, gt, gcount, gl, img_hw)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                hp.step(cls, reg, feats, gt, gcount, gl, img_hw)
        torch.cuda.current_stream().wait_stream(side)
        hps.append(hp); graphs.append(g); streams.append(side)
    return hps, graphs, streams


def run(D, prio=(0,), n=steps):
    hps, graphs, streams = build(D, prio)
    main = torch.cuda.current_stream()
    def go(m):
        start = torch.cuda.Event(enable_timing=True); end = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        start.record(main)
        for s in streams:
            s.wait_event(start)
        for i in range(m):
            with torch.cuda.stream(streams[i % D]):
                graphs[i % D].replay()
        for s in streams:
            main.wait_stream(s)
        end.record(main)
        torch.cuda.synchronize()
        return start.elapsed_time(end)
    go(2 * D)
    best = min(go(n) for _ in range(3))
    print("in flight %d  prio %-10s %7.1f us/step  (%6.0f images/s)" % (D, str(prio), best / n * 1e3, B * n / best * 1e3), flush=True)
    # the results of the last replay of every instance still equal a serial run's
    return hps


ref = run(1)
run(2)
run(2, (0, -1))
run(3)
run(4)
