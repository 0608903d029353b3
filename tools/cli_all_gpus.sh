#!/bin/bash
# The reference-facing seam at scale: the host mirror's `cli` (same flags as reference/nim/proof_input's cli) over a dataset of
# 16 slots x 16 GiB of the reference's fake data, once pinned to one GPU and once on every visible GPU (cdx_group_dataset_commit);
# the two input.json files must be identical.  usage: tools/cli_all_gpus.sh [out-dir]
set -e
cd "$(dirname "$0")/.."
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
ARGS="--field=bn254 --hash=poseidon2 --cellsize=2048 --blocksize=65536 --ncells=8388608 --nslots=16 --index=3 --nsamples=100 --seed=12345 --entropy=1234567 --depth=32 --maxslots=256"
CLI=codex-storage-proofs-circuits_b200/cli
t0=$(date +%s.%N); CODEX_COMMIT_GPUS=1 $CLI $ARGS --output=$OUT/cli_one_gpu.json > /dev/null; t1=$(date +%s.%N)
$CLI $ARGS --output=$OUT/cli_all_gpus.json > /dev/null; t2=$(date +%s.%N)
cmp $OUT/cli_one_gpu.json $OUT/cli_all_gpus.json && same=true || same=false
python - <<PY
import json
one, allg = $t1 - $t0, $t2 - $t1
gb = 16 * 16 * 2**30 / 1e9
print(json.dumps({"workload": "cli, 16 slots x 16 GiB of the reference's fake data (generated on the device), slot 3 sampled, 100 samples",
                  "one_gpu_wall_s": one, "all_gpus_wall_s": allg, "one_gpu_GB_per_s": gb / one, "all_gpus_GB_per_s": gb / allg,
                  "input_json_identical": "$same" == "true",
                  "note": "wall clock of the whole process: CUDA and NCCL initialisation, fake-data generation, commitment, proof input, JSON"}))
PY
rm -f $OUT/cli_one_gpu.json
