"""cuBLAS dgemm yard-stick (measurement only, never on the product path): SURVEY.md §7.2."""
import torch, time, json
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
out = {}
for (m, n, k) in [(8192, 8192, 8192), (5256, 104, 345600), (21024, 104, 345600 // 4), (104, 8192, 104)]:
    a = torch.randn(m, k, device=dev, dtype=torch.float64)
    b = torch.randn(k, n, device=dev, dtype=torch.float64)
    for _ in range(3): c = a @ b
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps): c = a @ b
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tf = 2.0 * m * n * k / ms * 1e-9
    out[f"{m}x{n}x{k}"] = {"ms": ms, "tflops": tf}
    print(f"dgemm {m}x{n}x{k}: {ms:.3f} ms {tf:.2f} TFLOP/s", flush=True)
    del a, b, c
json.dump(out, open("gpurun_out/dgemm_yardstick.json", "w"), indent=1)
