"""PCIe D2H probe (developer tool): contiguous vs the 2-D pattern phf_am_single_run_host uses (cudaMemcpy2DAsync)."""
import time
import torch
from cuda.bindings import runtime as rt

n, rows, w = 13440, 2000, 4
dev = torch.empty((n, rows, w), dtype=torch.float64, device="cuda")
host = torch.empty((n, rows, w), dtype=torch.float64).pin_memory()
stream = torch.cuda.current_stream().cuda_stream
gb = dev.numel() * 8 / 1e9


def t(fn, reps=4):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


print("contiguous D2H   %.1f GB/s" % (gb / t(lambda: host.copy_(dev, non_blocking=True))))
for seg in (1, 4, 8, 16, 32, 64):
    r = rows // seg
    d = torch.empty((n, r, w), dtype=torch.float64, device="cuda")

    def f():
        for k in range(seg):
            err, = rt.cudaMemcpy2DAsync(host.data_ptr() + k * r * w * 8, rows * w * 8, d.data_ptr(), r * w * 8,
                                        r * w * 8, n, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost, stream)
            assert err == rt.cudaError_t.cudaSuccess, err
    print("cudaMemcpy2DAsync, %2d segments of %4d rows (%6d B runs): %.1f GB/s" % (seg, r, r * w * 8, seg * r * n * w * 8 / 1e9 / t(f)))
