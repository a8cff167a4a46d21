# K-wave build variants (blocks per SM / warps per block), timed through bench.py's kernel_ms
mkdir -p gpurun_out
cd phnrec_b200/csrc
for v in "4 4" "5 4" "6 4" "3 8" "10 2"; do
  set -- $v
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-ffp-contract=off,-Wall -Xptxas -v --fmad=false -DPHN_PAIR_MINB=$1 -DPHN_PAIR_WARPS=$2 -c k_wave.cu -o build/k_wave.o 2> /tmp/ptx.log || { tail -5 /tmp/ptx.log; continue; }
  grep -A2 "Compiling.*k_wave_pairILi8ELb1" /tmp/ptx.log | grep -E "Used|spill" | tr '\n' ' '
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libphnrec_b200.so build/*.o -lcudart_static -lcuda -lpthread -ldl -lrt
  (cd ../..; timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('MINB=$1 WARPS=$2', j['ms_per_step'], j['kernel_ms'])")
done
