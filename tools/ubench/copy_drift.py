"""Does the GPU's copy / fill bandwidth hold over a long back-to-back run?  (The 704-block render writes 1.07 GB per
0.2 ms for 150 ms; its sustain step time creeps from 0.211 to 0.232 ms.)  Same measurement as MEASURED_PEAKS.json's
`hbm_gbs` (b.copy_(a) over 1 Gi bf16 elements, read + write bytes), 400 copies back to back, plus a write-only fill."""
import torch
n = 1 << 30
a = torch.empty(n, dtype=torch.bfloat16, device="cuda"); b = torch.empty_like(a)
a.fill_(1.0); torch.cuda.synchronize()
for name, fn, nbytes in (("copy", lambda: b.copy_(a), 4.0 * n), ("fill", lambda: b.fill_(2.0), 2.0 * n)):
    for rep in range(2):
        K = 400
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
        ev[0].record()
        for i in range(K):
            fn(); ev[i + 1].record()
        torch.cuda.synchronize()
        ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(K)]
        t = 0.0; marks = []
        for i, m in enumerate(ms):
            t += m
            if i in (0, 9, 49, 99, 199, 299, 399): marks.append(f"#{i + 1} (t={t:.0f} ms) {nbytes / m / 1e6:.0f} GB/s")
        print(name, "pass", rep, "best", f"{nbytes / min(ms) / 1e6:.0f}", "GB/s;", "; ".join(marks), flush=True)
