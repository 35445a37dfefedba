"""PCIe probe for the end-to-end leg: host->device bandwidth of 537 MB copies (one field of the bench workload) from ordinary
pinned memory vs write-combined pinned memory (cudaHostAllocWriteCombined), alone and with a device->host copy running the other
way -- the e2e number of bench.py is bounded by exactly this (1.88 GB up + 0.81 GB down per step)."""
import ctypes
import time

import torch

rt = ctypes.CDLL("libcudart.so.12")
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
N = 8192 * 8192 * 2          # floats: one float2 field of 8192^2 cells = 537 MB


def host(flags: int) -> torch.Tensor:
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), N * 4, flags) == 0
    t = torch.frombuffer((ctypes.c_float * N).from_address(p.value), dtype=torch.float32)
    t[::1024] = 1.0
    return t


dev_a, dev_b = torch.empty(N, device="cuda"), torch.ones(N, device="cuda")
down = torch.empty(N, dtype=torch.float32).pin_memory()
s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()
for name, flags in (("torch pin_memory()", None), ("cudaHostAlloc default", 0), ("cudaHostAlloc write-combined", 4)):
    src = torch.empty(N, dtype=torch.float32).pin_memory() if flags is None else host(flags)
    print(name, "is_pinned:", src.is_pinned())
    for both in (False, True):
        for _ in range(2):
            dev_a.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 6
        for _ in range(reps):
            with torch.cuda.stream(s_up):
                dev_a.copy_(src, non_blocking=True)
            if both:
                with torch.cuda.stream(s_down):
                    down.copy_(dev_b, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"  H2D {'with concurrent D2H' if both else 'alone':20s}: {reps * N * 4 / dt / 1e9:6.1f} GB/s per direction")
