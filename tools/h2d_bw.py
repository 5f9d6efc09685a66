"""Pinned host <-> device copy bandwidth of this box (context for bench.py's e2e number): python tools/h2d_bw.py"""
import torch
dev = torch.device("cuda:0")


def rate(fn, mb, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return mb * 1.048576 * reps / e0.elapsed_time(e1)


for mb in (2, 22, 256):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    print(f"{mb:4d} MiB own pinned tensor (untouched): H2D {rate(lambda: d.copy_(h, non_blocking=True), mb):6.1f} GB/s   D2H {rate(lambda: h.copy_(d, non_blocking=True), mb):6.1f} GB/s")
    h.fill_(3)
    print(f"{mb:4d} MiB own pinned tensor (written by the CPU first): H2D {rate(lambda: d.copy_(h, non_blocking=True), mb):6.1f} GB/s")
big = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
big.fill_(1)
d = torch.empty(22 << 20, dtype=torch.uint8, device=dev)
for off in (0, 100 << 20):
    sl = big[off:off + (22 << 20)]
    print(f"  22 MiB slice at +{off >> 20} MiB of a 256 MiB pinned arena: H2D {rate(lambda: d.copy_(sl, non_blocking=True), 22):6.1f} GB/s")
hr = torch.empty(22 << 20, dtype=torch.uint8)
torch.cuda.cudart().cudaHostRegister(hr.data_ptr(), hr.numel(), 0)
hr.fill_(2)
print(f"  22 MiB cudaHostRegister'ed pageable tensor: H2D {rate(lambda: d.copy_(hr, non_blocking=True), 22):6.1f} GB/s")
