import os, sys, subprocess, json, tempfile, time
ROOT = "/root/repo"
sys.path.insert(0, ROOT)
staged = os.path.join(ROOT, "baseline", "_ref", "code")
def run(tag, env=None):
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "h.json")
        subprocess.run([sys.executable, os.path.join(ROOT, "tests", "e2e_prove_dropin.py"), "gpu", out, "hello"],
                       env=dict(os.environ, B2S_REFERENCE_DIR=staged, **(env or {})), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
        print(tag, json.load(open(out))["prove_seconds"], flush=True)
run("plain parent")
run("plain parent again")
import torch
from stark_brainfuck_b200 import Engine
eng = Engine(0)
x = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
run("parent holds a CUDA context")
h = torch.empty(1 << 26, dtype=torch.uint8).pin_memory()
run("parent holds pinned memory too")
os.sched_setaffinity(0, set(list(os.sched_getaffinity(0))[:4]))
run("parent pinned to 4 cores")
