"""HBM write-only / read-only / copy bandwidth probe (torch, CUDA events): the roofline denominator of MEASURED_PEAKS.json is a
COPY (read + write); a kernel that only writes (the C5 ensemble flowpipe) or only reads sees a different ceiling."""
import json
import torch

dev = torch.device("cuda", 0)
n = 1 << 30     # 8 GiB of float64
a = torch.empty(n, dtype=torch.float64, device=dev)
b = torch.empty(n, dtype=torch.float64, device=dev)


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e-3


t_w = timed(lambda: a.zero_())
t_f = timed(lambda: a.fill_(1.5))
t_c = timed(lambda: b.copy_(a))
t_r = timed(lambda: a.sum())
print(json.dumps({"write_memset_gbs": 8 * n / t_w / 1e9, "write_fill_gbs": 8 * n / t_f / 1e9, "copy_rw_gbs": 16 * n / t_c / 1e9,
                  "read_sum_gbs": 8 * n / t_r / 1e9}))
