"""Write-only and read-only HBM streams beside the copy figure of MEASURED_PEAKS.json: the materialising kernels write
16-32x what they read, so their ceiling is the write stream, not the copy."""
import torch

dev = torch.device("cuda", 0)
n = 8 << 30
x = torch.empty(n, dtype=torch.uint8, device=dev)
y = torch.empty(n, dtype=torch.uint8, device=dev)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


ms = timed(lambda: x.zero_())
print(f"memset 8 GiB      : {ms:.2f} ms  {n / ms / 1e6:.0f} GB/s written")
ms = timed(lambda: x.view(torch.int64).fill_(0x0123456789))
print(f"fill_ 8 GiB (i64) : {ms:.2f} ms  {n / ms / 1e6:.0f} GB/s written")
ms = timed(lambda: y.copy_(x))
print(f"copy 8 GiB        : {ms:.2f} ms  {2 * n / ms / 1e6:.0f} GB/s read + written")
ms = timed(lambda: x.view(torch.int64).sum())
print(f"sum 8 GiB (i64)   : {ms:.2f} ms  {n / ms / 1e6:.0f} GB/s read")
