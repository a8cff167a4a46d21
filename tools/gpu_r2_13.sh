#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 120 python tools/e2e_timeline.py device 60 $O/r2m_tl_dev60.txt 2>&1 | tail -3
TL_FLUSH=1 timeout 120 python tools/e2e_timeline.py device 60 $O/r2m_tl_dev60f.txt 2>&1 | tail -3
PHNREC_VIT_INLINE=1 TL_FLUSH=1 timeout 120 python tools/e2e_timeline.py device 60 $O/r2m_tl_dev60fi.txt 2>&1 | tail -3
