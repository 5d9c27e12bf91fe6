"""Pure-store bandwidth on this GPU (context for the execute kernel's roofline; not a bench number of the library)."""
import ctypes as C
from pathlib import Path

import torch

lib = C.CDLL(str(Path(__file__).resolve().parent / "libfill.so"))
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return a.elapsed_time(e) / n * 1e-3


big = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
src = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
t = timeit(lambda: big.zero_())
print(f"torch zero_ 1 GiB: {big.numel() / t / 1e9:.0f} GB/s")
t = timeit(lambda: big.copy_(src))
print(f"torch copy 1 GiB: {2 * big.numel() / t / 1e9:.0f} GB/s (read+write), {big.numel() / t / 1e9:.0f} GB/s written")
for width in (16, 32):
    for grid, block in ((148 * 8, 128), (148 * 16, 128), (148 * 8, 256), (148 * 4, 512), (148 * 32, 128)):
        t = timeit(lambda: lib.fill_launch(C.c_void_p(big.data_ptr()), C.c_int64(big.numel()), width, grid, block, C.c_void_p(st)))
        print(f"grid-stride fill 1 GiB  st.{width * 8:3d}  grid {grid:5d} x {block:3d}: {big.numel() / t / 1e9:.0f} GB/s")
# the execute kernel's shape: 33.5 MB per launch, ring of 8 buffers (268 MB > L2), one CTA per 32 KiB tile
ring = [torch.empty(64 * 131072 * 4, dtype=torch.uint8, device=dev) for _ in range(8)]
for width in (16, 32):
    for tile, block in ((32768, 128), (16384, 128), (8192, 128), (65536, 256), (32768, 256)):
        i = [0]

        def f():
            b = ring[i[0] % 8]
            i[0] += 1
            lib.fill_tiled_launch(C.c_void_p(b.data_ptr()), C.c_int64(b.numel()), C.c_int64(tile), width, block, C.c_void_p(st))

        t = timeit(f, n=200, warm=20)
        print(f"tiled fill 33.5 MB x ring 8  st.{width * 8:3d}  tile {tile:6d} B x {block:3d} thr ({ring[0].numel() // tile} CTAs): "
              f"{t * 1e6:.2f} us/launch, {ring[0].numel() / t / 1e9:.0f} GB/s")
# lane-blocked stores (every lane writes 64 / 128 / 256 contiguous bytes): 2.68 GB like a 20-batch track ring
bigf = torch.empty(5 * (1 << 29), dtype=torch.uint8, device=dev)
for lb in (64, 128, 256):
    for tile, block in ((32768, 256), (32768, 128), (65536, 256)):
        t = timeit(lambda: lib.fill_blocked_launch(C.c_void_p(bigf.data_ptr()), C.c_int64(bigf.numel()), C.c_int64(tile), lb, block,
                                                   C.c_void_p(st)), n=10, warm=3)
        print(f"lane-blocked fill 2.68 GB  {lb:3d} B per lane  tile {tile} B x {block} thr: {bigf.numel() / t / 1e9:.0f} GB/s")
for tile, block in ((32768, 256), (32768, 128)):
    t = timeit(lambda: lib.fill_tiled_launch(C.c_void_p(bigf.data_ptr()), C.c_int64(bigf.numel()), C.c_int64(tile), 32, block,
                                             C.c_void_p(st)), n=10, warm=3)
    print(f"coalesced tiled fill 2.68 GB  st.256  tile {tile} B x {block} thr: {bigf.numel() / t / 1e9:.0f} GB/s")
i = [0]


def g():
    b = ring[i[0] % 8]
    i[0] += 1
    b.zero_()


t = timeit(g, n=200, warm=20)
print(f"torch zero_ 33.5 MB x ring 8: {t * 1e6:.2f} us/launch, {ring[0].numel() / t / 1e9:.0f} GB/s")
