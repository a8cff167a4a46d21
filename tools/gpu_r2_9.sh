#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 120 python tools/e2e_timeline.py sync 6 $O/r2i_tl_sync.txt 2>&1 | tail -9
PHNREC_NO_ARENA=1 timeout 120 python tools/e2e_timeline.py sync 6 $O/r2i_tl_sync_noarena.txt 2>&1 | tail -9
PHNREC_NO_ARENA=1 timeout 120 python tools/e2e_timeline.py async 8 $O/r2i_tl_async_noarena.txt 2>&1 | tail -11
