"""Independent 2^24 transforms issued alternately on two streams vs one stream (development aid)."""
import sys, torch
sys.path.insert(0, ".")
from toyni_b200 import device as D
from toyni_b200.lib import P
n = 1 << 24
bufs = [torch.randint(0, P, (n,), dtype=torch.int32, device="cuda") for _ in range(4)]
for nstreams in (1, 2, 3):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    for i in range(8):
        with torch.cuda.stream(streams[i % nstreams]):
            D.ntt_(bufs[i % 4])
    torch.cuda.synchronize()
    reps = 400
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams:
        s.wait_event(e0)
    for i in range(reps):
        with torch.cuda.stream(streams[i % nstreams]):
            D.ntt_(bufs[i % 4])
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    print(nstreams, "streams:", round(e0.elapsed_time(e1) * 1000 / reps, 2), "us per transform", flush=True)
