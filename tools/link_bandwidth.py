"""What the host <-> device link delivers for the e2e step's copies (pinned buffers, cudaMemcpyAsync through torch):
one direction alone, both directions at once, as 1 / 4 / 16 slices on their own streams.  61 MB each way = the q and
q_bar of 4096 pillow designs."""
import time

import torch

N = 4096 * 1860
dev = torch.device("cuda", 0)
h_in = torch.zeros(N, dtype=torch.float64).pin_memory()
h_out = torch.zeros(N, dtype=torch.float64).pin_memory()
d_in = torch.zeros(N, dtype=torch.float64, device=dev)
d_out = torch.zeros(N, dtype=torch.float64, device=dev)


h_small = torch.zeros(4096, dtype=torch.float64).pin_memory()
d_small = torch.zeros(4096, dtype=torch.float64, device=dev)
STATS = {}


def run(chunks, h2d, d2h, reps=20, small=False):
    streams = [torch.cuda.Stream(device=dev) for _ in range(chunks)]
    n = N // chunks
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for c, st in enumerate(streams):
            with torch.cuda.stream(st):
                if h2d:
                    d_in[c * n:(c + 1) * n].copy_(h_in[c * n:(c + 1) * n], non_blocking=True)
                if small:
                    m = 4096 // chunks
                    h_small[c * m:(c + 1) * m].copy_(d_small[c * m:(c + 1) * m], non_blocking=True)
                if d2h:
                    h_out[c * n:(c + 1) * n].copy_(d_out[c * n:(c + 1) * n], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = min(best, dt)
        STATS.setdefault((chunks, h2d, d2h, small), []).append(dt)
    return best


for chunks in (1, 4, 16):
    a, b, c = run(chunks, True, False), run(chunks, False, True), run(chunks, True, True)
    gb = N * 8 / 1e9
    print(f"chunks={chunks:2d}: H2D alone {gb / a:5.1f} GB/s ({a * 1e3:.3f} ms)  D2H alone {gb / b:5.1f} GB/s ({b * 1e3:.3f} ms)  "
          f"both at once {gb / c:5.1f} GB/s each way ({c * 1e3:.3f} ms)", flush=True)

import statistics
for small in (False, True):
    b = run(16, True, True, small=small)
    ts = STATS[(16, True, True, small)][-20:]
    print(f"16 slices, both ways, small loss copy per slice={small}: best {b * 1e3:.3f} ms  median {statistics.median(ts) * 1e3:.3f} ms  mean {statistics.mean(ts) * 1e3:.3f} ms")
