# A/B of two library builds on the same box: phnrec_b200/lib_old (PHNREC_B200_LIB) against the in-tree build
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_tensor_core.py -x -q 2>&1 | tail -3
for rep in 1 2; do
for lib in old new; do
  if [ $lib = old ]; then export PHNREC_B200_LIB=/root/repo/phnrec_b200/lib_old/libphnrec_b200.so; else unset PHNREC_B200_LIB; fi
  timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('$lib', round(j['ms_per_step'],3), 'e2e', round(j['e2e']['ms_per_step'],3), j['kernel_ms'])"
done; done
