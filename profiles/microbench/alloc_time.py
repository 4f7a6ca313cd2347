import time, torch
torch.cuda.init(); torch.zeros(1, device="cuda")
def t(n_bytes, reps):
    torch.cuda.synchronize(); keep=[]; t0=time.perf_counter()
    for _ in range(reps): keep.append(torch.empty(n_bytes, dtype=torch.uint8, device="cuda"))
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/reps*1e3
    return dt, keep
for mb in (8, 24, 64, 128, 264):
    dt, keep = t(mb<<20, 20)
    print(f"fresh {mb} MB: {dt:.3f} ms per torch.empty")
    del keep
    dt2, keep = t(mb<<20, 20)
    print(f"cached {mb} MB: {dt2:.3f} ms")
    del keep; torch.cuda.empty_cache()
