#!/usr/bin/env python3
"""Markdown table of the round-2 bench lines under profiles/ (the table in BASELINE.md section 4 is this script's output)."""
import json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
rows = []
base = None
for n in (1, 2, 4, 8):
    f = os.path.join(P, f"r2_bench_n{n}.json")
    if not os.path.exists(f):
        continue
    d = json.load(open(f))
    if n == 1:
        base = d
    eff = d["value"] / (n * base["value"]) if base else float("nan")
    eeff = d["e2e"]["value"] / (n * base["e2e"]["value"]) if base else float("nan")
    rows.append(f"| weak scaling, 10 GiB per GPU of one slot (config 3 at N = 1) | {n} | {d['config']['bytes_per_step']:,} | {d['ms_per_step'] / 1e3:.4f} | "
                f"{d['value']:.2f} ({100 * eff:.1f} %) | {d['e2e']['value']:.2f} ({100 * eeff:.1f} %) | {d['perms_per_s']:.3e} | {d['roofline']['frac']:.3f} "
                f"({d['roofline']['kernel_ms']:.1f} ms) | root identical on all ranks" + ("" if n == 1 else "; == whole slot on one GPU: " + str(d["sharded_root_check"]["equals_sharded_root"])) + " |")
    if "config4_strong" in d:
        c = d["config4_strong"]
        rows.append(f"| 4. ONE 100 GiB slot, strong scaling | {n} | 107,374,182,400 | {c['ms_per_step'] / 1e3:.4f} | {c['GB_per_s']:.2f} "
                    f"({100 * c['GB_per_s'] / (n * base['value']):.1f} %) | — | {1835008002 / (c['ms_per_step'] / 1e3):.3e} | — | root == `0x2df82ef9…894b` (one GPU, round 1): {c['root_equals_known_single_gpu_root']} |")
    if "config5" in d:
        c = d["config5"]
        rows.append(f"| 5. dataset, 250 slots 0.25-25 GiB + 100 paths | {n} | {c['bytes']:,} | {c['commit_s']:.2f} | {c['GB_per_s']:.2f} "
                    f"({100 * c['GB_per_s'] / (n * base['value']):.1f} %) | — | — | — | all 100 paths reconstruct the slot root: {c['all_100_paths_reconstruct_slot_root']}; "
                    f"rank loads max/min {c['per_rank_bytes_max_over_min']:.4f} |")
print("| Config | GPUs | bytes | seconds | GB/s resident (vs N x one GPU) | GB/s end to end from host buffers | perms/s | roofline frac (cell kernel) | parity |")
print("|---|---|---|---|---|---|---|---|---|")
print("\n".join(rows))
if base:
    s = base["small_slots"]
    print(f"\nN = 1 extras: 2^20-permutation batch {base['perm_batch_2^20']['ms']:.3f} ms = {base['perm_batch_2^20']['perms_per_s']:.3e} perms/s; "
          f"config 1 (11 x 4 MiB fake slots + dataset root) {s['config1_11x4MiB_fake']['batched_ms']:.1f} ms batched vs {s['config1_11x4MiB_fake']['one_by_one_ms']:.1f} ms one by one; "
          f"1000 x 4 MiB slots {s['batch_1000x4MiB']['GB_per_s']:.2f} GB/s batched vs {s['batch_1000x4MiB']['one_by_one_GB_per_s']:.2f} one by one; "
          f"CPU baseline {base['cpu_baseline']['value']:.3f} GB/s on {base['cpu_baseline']['cores']} threads.")
