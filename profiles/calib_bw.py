"""What the memory system gives to the three access mixes of this path, measured with library ops (calibration only):
copy (read+write, the MEASURED_PEAKS.json mix), pure write (the slice kernel's mix: ~1 % reads) and pure read (the merge
kernel's mix: ~3 % writes). Bytes counted once per element moved; CUDA-event timed, 9 GB working set (>> L2)."""
import json, torch
dev = torch.device("cuda", 0)
n = 9 * 1024**3 // 4
a = torch.empty(n, dtype=torch.float32, device=dev); b = torch.empty(n, dtype=torch.float32, device=dev)
a.fill_(1.0); b.fill_(2.0)
def timed(fn, steps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / steps
out = {}
ms = timed(lambda: b.copy_(a)); out["copy_gbs"] = 2 * n * 4 / ms / 1e6
ms = timed(lambda: a.zero_()); out["memset_write_gbs"] = n * 4 / ms / 1e6
ms = timed(lambda: a.fill_(3.0)); out["fill_write_gbs"] = n * 4 / ms / 1e6
ms = timed(lambda: a.sum()); out["sum_read_gbs"] = n * 4 / ms / 1e6
ms = timed(lambda: torch.add(a, b, out=b)); out["add_2r1w_gbs"] = 3 * n * 4 / ms / 1e6
print(json.dumps({k: round(v, 1) for k, v in out.items()}))
