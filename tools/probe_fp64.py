"""Probe: cuBLAS DGEMM (torch.matmul f64) throughput + host CPU info on the B200 box."""
import os, time, torch, json
print("cpu_count", os.cpu_count(), "torch threads", torch.get_num_threads())
print(torch.cuda.get_device_name(0))
res = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(3): c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[f"dgemm_{n}"] = 2 * n**3 / (best * 1e-3) * 1e-12
    print(n, "DGEMM TFLOP/s", res[f"dgemm_{n}"], "ms", best)
# sustained
n = 8192
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(60): c = a @ b
e1.record(); torch.cuda.synchronize()
res["dgemm_8192_sustained"] = 60 * 2 * n**3 / (e0.elapsed_time(e1) * 1e-3) * 1e-12
print("sustained", res["dgemm_8192_sustained"])
# cholesky f64 via cusolver for reference
for n in (4096, 8192):
    x = torch.randn(n, n, dtype=torch.float64, device="cuda"); k = x @ x.T + n * torch.eye(n, dtype=torch.float64, device="cuda")
    torch.linalg.cholesky(k); torch.cuda.synchronize()
    t = time.perf_counter(); L = torch.linalg.cholesky(k); torch.cuda.synchronize(); dt = time.perf_counter() - t
    res[f"potrf_{n}_ms"] = dt * 1e3
    print("cusolver potrf", n, dt * 1e3, "ms")
json.dump(res, open("gpurun_out/probe_fp64.json", "w"))
