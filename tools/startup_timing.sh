set -e
root=/root/repo; ref=$root/oracle/_ref; work=$(mktemp -d /dev/shm/vt.XXXXXX); trap 'rm -rf "$work"' EXIT
"$root/oracle/synth" cfg -o "$work/c" -L 1000000 -n 1000 -r 1000000 -s 1 >/dev/null
"$ref/snp-pattern-gen" -k 21 -f "$work/c.fa" -b "$work/c.bed" -o "$work/p.txt" 2>/dev/null
for t in 1 4; do
s=$(date +%s%N); VAFGPU_TIMING=1 "$root/kmer-cnt_b200/vaf-counter" -k 21 -t $t -v -p "$work/p.txt" -o "$work/o.vaf" "$work/c.fq" 2> "$work/err"; e=$(date +%s%N)
echo "-t $t wall $(( (e - s) / 1000000 )) ms"; grep "vafgpu\]\|Total runtime\|K-mer map\|K-mer counting\|Pattern loading" "$work/err"
done
s=$(date +%s%N); python3 - <<'PY'
import ctypes
ctypes.CDLL("libcuda.so.1").cuInit(0)
PY
e=$(date +%s%N); echo "python + cuInit only: $(( (e - s) / 1000000 )) ms"
cat > "$work/t.cu" <<'CU'
#include <cuda_runtime.h>
#include <cstdio>
#include <chrono>
int main(){auto t0=std::chrono::steady_clock::now(); cudaFree(0); auto t1=std::chrono::steady_clock::now(); void*p; cudaMallocHost(&p,30<<20); auto t2=std::chrono::steady_clock::now(); cudaMalloc(&p,(size_t)1<<30); auto t3=std::chrono::steady_clock::now();
printf("context %.1f ms, pinned 30MB %.1f ms, malloc 1GB %.1f ms\n", std::chrono::duration<double,std::milli>(t1-t0).count(), std::chrono::duration<double,std::milli>(t2-t1).count(), std::chrono::duration<double,std::milli>(t3-t2).count()); return 0;}
CU
nvcc -o "$work/t" "$work/t.cu" 2>/dev/null && s=$(date +%s%N) && "$work/t" && e=$(date +%s%N) && echo "minimal CUDA program wall $(( (e - s) / 1000000 )) ms"
