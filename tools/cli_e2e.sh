#!/bin/bash
# End-to-end check of the command line on BASELINE config 0 (and a 10x larger read set):
# snp-pattern-gen -k 21 on a synthetic 1 Mb FASTA + 1k-SNP BED, then vaf-counter -k 21 on
# synthetic 150 bp reads -- the unmodified reference (oracle/_ref) against this repo's CLI,
# same files, byte comparison of the .vaf, wall clock of the whole process and the tools' own
# "Speed:" lines.  Usage: tools/cli_e2e.sh [reads ...]   (default 1000000 10000000)
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
ref=$root/oracle/_ref
work=$(mktemp -d /dev/shm/cli_e2e.XXXXXX 2>/dev/null || mktemp -d)
trap 'rm -rf "$work"' EXIT
ncpu=$(nproc)
for reads in ${@:-1000000 10000000}; do
	"$root/oracle/synth" cfg -o "$work/c" -L 1000000 -n 1000 -r "$reads" -s 1 >/dev/null
	"$ref/snp-pattern-gen" -k 21 -f "$work/c.fa" -b "$work/c.bed" -o "$work/p.txt" 2>/dev/null
	echo "== $reads reads x 150 bp, $(wc -l < "$work/p.txt") patterns, FASTQ $(du -h "$work/c.fq" | cut -f1), host has $ncpu cores"
	for t in 1 4 $ncpu; do
		s=$(date +%s%N)
		"$ref/vaf-counter" -k 21 -t $t -v -p "$work/p.txt" -o "$work/ref$t.vaf" "$work/c.fq" 2> "$work/ref$t.err"
		e=$(date +%s%N)
		echo "reference -t $t: wall $(( (e - s) / 1000000 )) ms; $(grep 'Speed:' "$work/ref$t.err" | tr -s ' ')"
	done
	for t in 1 4 $ncpu; do
		s=$(date +%s%N)
		"$root/kmer-cnt_b200/vaf-counter" -k 21 -t $t -v -p "$work/p.txt" -o "$work/gpu$t.vaf" "$work/c.fq" 2> "$work/gpu$t.err"
		e=$(date +%s%N)
		cmp -s "$work/gpu$t.vaf" "$work/ref1.vaf" && same=identical || same=DIFFERENT
		echo "this repo -t $t: wall $(( (e - s) / 1000000 )) ms; $(grep 'Speed:' "$work/gpu$t.err" | tr -s ' '); $(grep 'Kernel time' "$work/gpu$t.err" | tr -s ' '); .vaf $same"
	done
	cmp -s "$work/ref1.vaf" "$work/ref4.vaf" || echo "reference -t 1 and -t 4 differ?!"
	head -3 "$work/ref1.vaf"
done
