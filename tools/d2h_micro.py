"""PCIe D2H microbenchmark: one step's observation block (c2: 2.4 MB) copied with 1..4 streams / chunks."""
import time
import torch

nbytes = 4096 * 14 * 14 * 3
dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
streams = [torch.cuda.Stream() for _ in range(4)]
for nchunk, nstream in [(1, 1), (2, 1), (2, 2), (4, 2), (4, 4), (8, 4)]:
    step = nbytes // nchunk
    def once():
        for c in range(nchunk):
            with torch.cuda.stream(streams[c % nstream]):
                host[c * step:(c + 1) * step].copy_(dev[c * step:(c + 1) * step], non_blocking=True)
        for s in streams[:nstream]:
            s.synchronize()
    for _ in range(20):
        once()
    t0 = time.perf_counter()
    R = 300
    for _ in range(R):
        once()
    dt = (time.perf_counter() - t0) / R
    print(f"chunks={nchunk} streams={nstream}: {dt * 1e6:.1f} us  {nbytes / dt / 1e9:.1f} GB/s", flush=True)
# big copy for reference
big_d = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); big_h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter(); big_h.copy_(big_d, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"256 MiB: {(256 << 20) / dt / 1e9:.1f} GB/s")
