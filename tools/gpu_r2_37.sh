#!/bin/bash
# feasibility of a decoder that co-resides with the MLP kernel: MLP with 27 KB less shared memory, decoder capped at 96 registers
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
L=$PWD/phnrec_b200/lib
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    j=json.load(open(f"gpurun_out/{n}.json")); print(f"{n:18s}", round(j["ms_per_step"],3), "ms e2e", round(j["e2e"]["ms_per_step"],3), [(k["kernel"], k["ms"]) for k in j["roofline"]["kernels"]])
except Exception as e: print(n, "ERR", e, open(f"gpurun_out/{n}.err").read()[-1500:])
PY
}
B="--steps 20 --warmup 3 --no-cpu-baseline --no-parity --profile-seconds 1"
timeout 120 python bench.py $B > $O/r2U_base.json 2> $O/r2U_base.err; show r2U_base
PHNREC_TC_SMEM=204800 timeout 120 python bench.py $B > $O/r2U_smem200.json 2> $O/r2U_smem200.err; show r2U_smem200
PHNREC_TC_SMEM=196608 timeout 120 python bench.py $B > $O/r2U_smem192.json 2> $O/r2U_smem192.err; show r2U_smem192
PHNREC_B200_LIB=$L/libphnrec_b200_vit96.so timeout 120 python bench.py $B > $O/r2U_vit96.json 2> $O/r2U_vit96.err; show r2U_vit96
