#!/usr/bin/env python3
"""Start-up and commit time of the all-GPUs-of-one-process group (cdx_group_*) against one context, on the reference's fake
data; prints one JSON line.  usage: group_timing.py [n_slots = 16] [GiB per slot = 16]"""
import ctypes as C, importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("codex-storage-proofs-circuits_b200")
capi = pkg.capi
lib = pkg.load_library()
n_slots = int(sys.argv[1]) if len(sys.argv) > 1 else 16
gib = float(sys.argv[2]) if len(sys.argv) > 2 else 16.0
n_bytes = int(gib * (1 << 30)) // 65536 * 65536
kind = capi.SRC_FAKE if os.environ.get("KIND", "fake") == "fake" else capi.SRC_SYNTHETIC
descs, keep = capi.make_descs([(kind, 1000 + k, n_bytes) for k in range(n_slots)])
out = {"n_slots": n_slots, "gib_per_slot": gib, "n_gpus": lib.cdx_device_count(), "source": "fake" if kind == capi.SRC_FAKE else "synthetic"}
t0 = time.perf_counter()
ctx = pkg.Context(0)
out["one_ctx_create_s"] = time.perf_counter() - t0
t0 = time.perf_counter()
ds = ctx.dataset_commit(None, [(kind, 1000 + k, n_bytes) for k in range(n_slots)])
out["one_gpu_commit_s"] = time.perf_counter() - t0
root1 = ds.root
ds.free()
t0 = time.perf_counter()
g = C.c_void_p()
assert lib.cdx_group_create(None, 0, C.byref(g)) == 0
out["group_create_s"] = time.perf_counter() - t0
n = lib.cdx_group_size(g)
for rep in range(2):
    handles = (C.c_void_p * n)()
    t0 = time.perf_counter()
    rc = lib.cdx_group_dataset_commit(g, descs, n_slots, 2048, 65536, -1, handles)
    dt = time.perf_counter() - t0
    assert rc == 0, lib.cdx_group_last_error(g)
    out[f"group_commit_s_{rep}"] = dt
    root = C.create_string_buffer(32)
    lib.cdx_dataset_root(handles[0], root)
    assert int.from_bytes(root.raw, "little") == root1
    lib.cdx_group_datasets_free(g, handles)
total = n_slots * n_bytes
out["one_gpu_GB_per_s"] = total / out["one_gpu_commit_s"] / 1e9
out["group_GB_per_s"] = total / out["group_commit_s_1"] / 1e9
lib.cdx_group_destroy(g)
print(json.dumps(out))
