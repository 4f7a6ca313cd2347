import sys, time, runpy, torch
sys.path.insert(0, "/root/repo")
from stark_brainfuck_b200 import engine as E
log = []
orig = E.Engine.alloc
def alloc(self, shape, dtype=torch.int64, zero=False):
    t0 = time.perf_counter(); r = orig(self, shape, dtype, zero); dt = (time.perf_counter() - t0) * 1e3
    log.append((dt, tuple(shape), torch.cuda.memory_reserved() >> 20))
    return r
E.Engine.alloc = alloc
sys.argv = ["e2e", "gpu", "gpurun_out/r02bi.json", "+++++++++++[>+++++++++++[>+++++++++++[>+<-]<-]<-]"]
try:
    runpy.run_path("tests/e2e_prove_dropin.py", run_name="__main__")
finally:
    print("allocs", len(log), "total ms", sum(d for d, _, _ in log))
    for d, s, r in sorted(log, reverse=True)[:25]:
        print(f"{d:8.2f} ms  {s}  reserved {r} MB")
