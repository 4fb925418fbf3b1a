"""Development: HBM write-only and read-only rates against the copy rate of MEASURED_PEAKS.json (torch fill_ / sum / copy_ on
16 GiB buffers, CUDA events) -- the stash write of the forward E-step kernel is a write-only stream, its read by the
backward kernel a read-only one."""
import torch
n = (16 << 30) // 4
x = torch.empty(n, dtype=torch.float32, device="cuda")
y = torch.empty(n, dtype=torch.float32, device="cuda")
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t = timed(lambda: x.fill_(1.0)); print("write-only  %.1f GB/s" % (n * 4 / t / 1e6))
t = timed(lambda: y.copy_(x)); print("copy        %.1f GB/s (read + write bytes)" % (2 * n * 4 / t / 1e6))
t = timed(lambda: x.sum()); print("read-only   %.1f GB/s" % (n * 4 / t / 1e6))
