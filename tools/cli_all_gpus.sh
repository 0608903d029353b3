#!/bin/bash
# The reference-facing seam at scale: the host mirror's `cli` (same flags as reference/nim/proof_input's cli) over a dataset of
# 16 slots x 16 GiB of the reference's fake data, once pinned to one GPU and once on every visible GPU (cdx_group_dataset_commit);
# the two input.json files must be identical.  Reports the wall clock of each process and, from CODEX_HOST_TRACE, the time
# generateProofInputBN254 needed up to "dataset committed" (process start-up -- CUDA context creation takes 2-5 s per fresh
# box -- and teardown are outside it).  usage: tools/cli_all_gpus.sh [out-dir]
set -e
cd "$(dirname "$0")/.."
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
ARGS="--field=bn254 --hash=poseidon2 --cellsize=2048 --blocksize=65536 --ncells=8388608 --nslots=16 --index=3 --nsamples=100 --seed=12345 --entropy=1234567 --depth=32 --maxslots=256"
CLI=codex-storage-proofs-circuits_b200/cli
export CODEX_HOST_TRACE=1
t0=$(date +%s.%N); CODEX_COMMIT_GPUS=1 $CLI $ARGS --output=$OUT/cli_one_gpu.json > /dev/null 2> $OUT/cli_one_gpu.trace; t1=$(date +%s.%N)
$CLI $ARGS --output=$OUT/cli_all_gpus.json > /dev/null 2> $OUT/cli_all_gpus.trace; t2=$(date +%s.%N)
cmp $OUT/cli_one_gpu.json $OUT/cli_all_gpus.json && same=true || same=false
python - <<PY
import json, re
def milestone(path, what):
    for line in open(path):
        m = re.match(r"\[trace\]\s+([0-9.]+) s\s+(.*)", line)
        if m and what in m.group(2):
            return float(m.group(1))
one, allg = $t1 - $t0, $t2 - $t1
gb = 16 * 16 * 2**30 / 1e9
c1 = milestone("$OUT/cli_one_gpu.trace", "dataset committed")
g0 = milestone("$OUT/cli_all_gpus.trace", "group of all GPUs created")
c8 = milestone("$OUT/cli_all_gpus.trace", "dataset committed")
print(json.dumps({"workload": "cli, 16 slots x 16 GiB of the reference's fake data (generated on the device), slot 3 sampled, 100 samples",
                  "one_gpu": {"process_wall_s": one, "dataset_committed_s": c1, "GB_per_s": gb / c1},
                  "all_gpus": {"process_wall_s": allg, "group_created_s": g0, "dataset_committed_s": c8, "commit_only_s": c8 - g0, "GB_per_s_commit_only": gb / (c8 - g0)},
                  "input_json_identical": "$same" == "true"}))
PY
rm -f $OUT/cli_one_gpu.json $OUT/cli_all_gpus.json
