#!/bin/bash
# End-to-end check of the yak-count command line: the unmodified reference (oracle/_ref/yak-count)
# against this repo's CLI on the same synthetic FASTQ (150 bp reads from a 50 Mb genome, 1 %
# substitutions, 0.5 % N), without and with the Bloom filter (-b 30), byte comparison of the
# 1023 histogram lines, wall clock of the whole process.
# Usage: tools/yak_cli_e2e.sh [reads ...]   (default 1000000 5000000)
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
ref=$root/oracle/_ref
work=$(mktemp -d /dev/shm/yak_e2e.XXXXXX 2>/dev/null || mktemp -d)
trap 'rm -rf "$work"' EXIT
ncpu=$(nproc)
# the first CUDA process on a fresh box pays for the driver paging in: not part of either tool
"$root/oracle/synth" cfg -o "$work/w" -L 100000 -n 2 -r 1000 -s 3 >/dev/null 2>&1
"$root/kmer-cnt_b200/yak-count" -k 31 -t 1 "$work/w.fq" >/dev/null 2>&1 || true
for reads in ${@:-1000000 5000000}; do
	"$root/oracle/synth" cfg -o "$work/c" -L 50000000 -n 10 -r "$reads" -e 0.01 -N 0.005 -s 5 >/dev/null 2>&1
	echo "== $reads reads x 150 bp, k = 31, FASTQ $(du -h "$work/c.fq" | cut -f1), host has $ncpu cores"
	for b in 0 30; do
		s=$(date +%s%N)
		"$ref/yak-count" -k 31 -b $b -t $ncpu "$work/c.fq" > "$work/ref$b.hist" 2>/dev/null
		e=$(date +%s%N)
		echo "reference -b $b -t $ncpu: wall $(( (e - s) / 1000000 )) ms = $(( reads * 150 * 1000 / ((e - s) / 1000) / 1000 )) Mbases/s"
		for t in 1 $ncpu; do
			s=$(date +%s%N)
			KCGPU_TIMING=1 "$root/kmer-cnt_b200/yak-count" -k 31 -b $b -t $t "$work/c.fq" > "$work/gpu$b.$t.hist" 2> "$work/gpu$b.$t.err"
			e=$(date +%s%N)
			cmp -s "$work/gpu$b.$t.hist" "$work/ref$b.hist" && same=identical || same=DIFFERENT
			echo "this repo -b $b -t $t: wall $(( (e - s) / 1000000 )) ms = $(( reads * 150 * 1000 / ((e - s) / 1000) / 1000 )) Mbases/s; histogram $same; $(grep distinct "$work/gpu$b.$t.err")"
			grep '^\[yak-count\]' "$work/gpu$b.$t.err" | sed 's/^/    /'
		done
	done
	head -3 "$work/ref30.hist" | tr '\n' ' '; echo
done
