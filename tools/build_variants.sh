#!/bin/bash
# Development: build differently tuned copies of libvafgpu.so into tools/exp/variants/ (git-ignored;
# they travel to the GPU box) for tools/kvar.py to time side by side.
#   tools/build_variants.sh name:"-DVG_AHEAD=2 -DVG_PF_BYTES=8192" other:"..."
# `base:<git-rev>` builds the library as of that revision (the kernel a change is compared with).
set -euo pipefail
root=$(cd "$(dirname "$0")/.." && pwd)
out="$root/tools/exp/variants"
mkdir -p "$out"
build() { # dir name flags
	( cd "$1/kmer-cnt_b200" && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -diag-suppress 550 \
		-Xcompiler -fPIC $3 -shared -o "$out/$2.so" csrc/vafgpu_api.cu csrc/vafgpu_kernels.cu csrc/vafgpu_tables.cpp \
		csrc/vafgpu_pack.cpp csrc/kcgpu_api.cu csrc/kcgpu_kernels.cu -ldl && echo "built $2: $3" ) &
}
for spec in "$@"; do
	name=${spec%%:*}; flags=${spec#*:}
	if [ "$name" = base ]; then
		tmp=$(mktemp -d /tmp/vafbase.XXXX)
		git -C "$root" archive "$flags" kmer-cnt_b200 include | tar -x -C "$tmp"
		build "$tmp" "base_$flags" ""
	else
		build "$root" "$name" "$flags"
	fi
	while [ "$(jobs -r | wc -l)" -ge 6 ]; do sleep 1; done
done
wait
ls -la "$out"
