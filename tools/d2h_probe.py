"""How fast does this box read device memory back, and does it matter how the copies are issued? (one 1080p rgba16f image = 16.6 MB)"""
import json, time, torch
n, reps = 1920 * 1080 * 8, 32
dev = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(reps)]
host = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(reps)]
out = {}
for streams in (1, 2, 4):
    ss = [torch.cuda.Stream() for _ in range(streams)]
    torch.cuda.synchronize()
    for _ in range(2):
        t0 = time.perf_counter()
        for i in range(reps):
            with torch.cuda.stream(ss[i % streams]):
                host[i].copy_(dev[i], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    out[f"d2h_{streams}_streams_GBps"] = n * reps / dt / 1e9
# one image split in 4 chunks over 4 streams
ss = [torch.cuda.Stream() for _ in range(4)]
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(reps):
    for k in range(4):
        with torch.cuda.stream(ss[k]):
            host[i][k * n // 4:(k + 1) * n // 4].copy_(dev[i][k * n // 4:(k + 1) * n // 4], non_blocking=True)
torch.cuda.synchronize(); out["d2h_split4_GBps"] = n * reps / (time.perf_counter() - t0) / 1e9
big_d = torch.empty(n * reps, dtype=torch.uint8, device="cuda"); big_h = torch.empty(n * reps, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter(); big_h.copy_(big_d, non_blocking=True); torch.cuda.synchronize()
out["d2h_one_531MB_copy_GBps"] = n * reps / (time.perf_counter() - t0) / 1e9
torch.cuda.synchronize(); t0 = time.perf_counter(); big_d.copy_(big_h, non_blocking=True); torch.cuda.synchronize()
out["h2d_one_531MB_copy_GBps"] = n * reps / (time.perf_counter() - t0) / 1e9
print(json.dumps(out))
