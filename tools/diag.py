#!/usr/bin/env python3
"""diagnostics: H2D bandwidth, alloc/free cost, commit step time with and without nvidia-smi polling"""
import importlib, os, sys, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("codex-storage-proofs-circuits_b200")
ctx = pkg.Context(0)
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
n = int(gib * (1 << 30))
d = torch.empty(n, dtype=torch.uint8, device="cuda")
ctx.fill_synthetic_dev(0xC0DE, 0, n, d.data_ptr()); torch.cuda.synchronize()
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); h.copy_(d); torch.cuda.synchronize()
for _ in range(2):
    t0 = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"H2D pinned {gib} GiB: {n/dt/1e9:.1f} GB/s")
t0 = time.perf_counter(); x = torch.empty(340 << 20, dtype=torch.uint8, device="cuda"); torch.cuda.synchronize(); print("torch alloc 340MB ms", 1e3*(time.perf_counter()-t0))
def step():
    s = ctx.slot_commit_dev(d.data_ptr(), n, 2048, 65536); r = s.root; s.free(); return r
def step_host():
    s = ctx.slot_commit_host(h.numpy(), 2048, 65536); r = s.root; s.free(); return r
for name, fn in (("dev", step), ("host", step_host)):
    fn()
    for trial in range(2):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"{name} commit {gib} GiB: {1e3*dt:.1f} ms  {n/dt/1e9:.2f} GB/s")
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader", "-lms", "100"], stdout=subprocess.DEVNULL)
time.sleep(0.5)
for trial in range(2):
    t0 = time.perf_counter(); step(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"dev commit with nvidia-smi -lms 100 polling: {1e3*dt:.1f} ms  {n/dt/1e9:.2f} GB/s")
p.terminate()
t0 = time.perf_counter(); s = ctx.slot_commit_dev(d.data_ptr(), n, 2048, 65536); t1 = time.perf_counter(); r = s.root; t2 = time.perf_counter(); s.free(); t3 = time.perf_counter()
print(f"launch {1e3*(t1-t0):.1f} ms, root(sync) {1e3*(t2-t1):.1f} ms, free {1e3*(t3-t2):.1f} ms")
# pynvml in-process sampler perturbation test
import threading
try:
    import pynvml
    pynvml.nvmlInit()
    hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
    stop = [False]; samples = []
    def poll():
        while not stop[0]:
            samples.append((pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetCurrentClocksEventReasons(hnd)))
            time.sleep(0.1)
    th = threading.Thread(target=poll, daemon=True); th.start()
    for trial in range(4):
        t0 = time.perf_counter(); step(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"dev commit with pynvml polling: {1e3*dt:.1f} ms  {n/dt/1e9:.2f} GB/s")
    stop[0] = True; th.join()
    print("pynvml samples", len(samples), samples[:3], "max_sm", pynvml.nvmlDeviceGetMaxClockInfo(hnd, pynvml.NVML_CLOCK_SM))
except Exception as e:
    print("pynvml failed", repr(e))
for trial in range(3):
    t0 = time.perf_counter(); s = ctx.slot_commit_dev(d.data_ptr(), n, 2048, 65536); t1 = time.perf_counter(); r = s.root; t2 = time.perf_counter(); s.free(); t3 = time.perf_counter()
    print(f"launch {1e3*(t1-t0):.1f} ms, root(sync) {1e3*(t2-t1):.1f} ms, free {1e3*(t3-t2):.1f} ms")
