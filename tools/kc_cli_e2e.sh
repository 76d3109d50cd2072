#!/bin/bash
# End-to-end check of the kc-c4 command line: the unmodified reference (oracle/_ref/kc-c4)
# against this repo's CLI on the same synthetic FASTQ (150 bp reads from a 50 Mb genome, 1 %
# substitutions, 0.5 % N), byte comparison of the 255 histogram lines, wall clock of the whole
# process.  Usage: tools/kc_cli_e2e.sh [reads ...]   (default 1000000 10000000)
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
ref=$root/oracle/_ref
work=$(mktemp -d /dev/shm/kc_e2e.XXXXXX 2>/dev/null || mktemp -d)
trap 'rm -rf "$work"' EXIT
ncpu=$(nproc)
for reads in ${@:-1000000 10000000}; do
	"$root/oracle/synth" cfg -o "$work/c" -L 50000000 -n 10 -r "$reads" -e 0.01 -N 0.005 -s 5 >/dev/null 2>&1
	echo "== $reads reads x 150 bp, k = 31, FASTQ $(du -h "$work/c.fq" | cut -f1), host has $ncpu cores"
	for t in 1 4 $ncpu; do
		s=$(date +%s%N)
		"$ref/kc-c4" -k 31 -t $t "$work/c.fq" > "$work/ref$t.hist"
		e=$(date +%s%N)
		echo "reference -t $t: wall $(( (e - s) / 1000000 )) ms = $(( reads * 150 * 1000 / ((e - s) / 1000) / 1000 )) Mbases/s"
	done
	for t in 1 4 $ncpu; do
		s=$(date +%s%N)
		KCGPU_TIMING=1 "$root/kmer-cnt_b200/kc-c4" -k 31 -t $t "$work/c.fq" > "$work/gpu$t.hist" 2> "$work/gpu$t.err"
		e=$(date +%s%N)
		cmp -s "$work/gpu$t.hist" "$work/ref1.hist" && same=identical || same=DIFFERENT
		echo "this repo -t $t: wall $(( (e - s) / 1000000 )) ms = $(( reads * 150 * 1000 / ((e - s) / 1000) / 1000 )) Mbases/s; histogram $same"
		sed 's/^/    /' "$work/gpu$t.err"
	done
	head -3 "$work/ref1.hist" | tr '\n' ' '; echo
done
