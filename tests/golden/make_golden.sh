#!/bin/bash
# Regenerates the committed fixtures from the UNMODIFIED reference (needs /root/reference,
# i.e. this only runs in the build container).  Everything written here is an OUTPUT of the
# reference tools or an input we generated ourselves; no reference source is copied.
#   kat_vaf.tsv            known answers of the reference's own functions (oracle/ref_kat.c)
#   e2e_*/                 small inputs + the .vaf the reference vaf-counter wrote for them
#   kat_kc.tsv             known answers of kc-c4's hash64 / count_seq_buf (oracle/ref_kat_kc.c)
#   kc/<reads>.k<K>.hist   what the reference kc-c4 -k K prints for e2e_<reads>/reads.fq.gz
#   cfg2_patterns.txt.gz   snp-pattern-gen -k 21 over a synthetic hg38-length genome and the
#                          NGSCheckMate GRCh38 panel (made separately, see tools/make_cfg2_patterns.sh)
set -euo pipefail
here=$(cd "$(dirname "$0")" && pwd)
root=$(cd "$here/../.." && pwd)
make -s -C "$root/oracle" all
ref="$root/oracle/_ref"
synth="$root/oracle/synth"
"$ref/ref_kat" > "$here/kat_vaf.tsv" 2>/dev/null

mk() { # name k synth-args...
	local name=$1 k=$2; shift 2
	local d="$here/e2e_$name"; rm -rf "$d"; mkdir -p "$d"
	"$synth" cfg -o "$d/x" "$@"
	"$ref/snp-pattern-gen" -k "$k" -b "$d/x.bed" -f "$d/x.fa" -o "$d/patterns.txt" 2>/dev/null
	"$ref/vaf-counter" -k "$k" -t 2 -b 200000 -p "$d/patterns.txt" -o "$d/expected.vaf" "$d/x.fq" 2>/dev/null
	"$ref/vaf-counter-scalar" -k "$k" -t 1 -p "$d/patterns.txt" -o "$d/expected_scalar.vaf" "$d/x.fq" 2>/dev/null
	mv "$d/x.fq" "$d/reads.fq"; gzip -9n "$d/reads.fq"
	rm -f "$d/x.fa" "$d/x.bed"
	echo "$k" > "$d/k"
}
mk k21 21 -L 60000 -n 300 -r 4000 -l 150 -s 11
mk k15 15 -L 60000 -n 300 -r 4000 -l 150 -s 12 -N 0.02 -M 4
mk k31 31 -L 60000 -n 300 -r 4000 -l 150 -s 13 -N 0.05 -M 5
mk exotic 21 -L 60000 -n 300 -r 4000 -l 150 -j 40 -s 14 -x 0.01

# counting mode: known answers of kc-c4's own functions, and the histogram the reference kc-c4
# prints for the read sets above
"$ref/ref_kat_kc" > "$here/kat_kc.tsv"
rm -rf "$here/kc"; mkdir -p "$here/kc"
for name in k21 k15 k31 exotic; do
	for k in 5 15 21 28 31; do
		"$ref/kc-c4" -k "$k" -t 2 "$here/e2e_$name/reads.fq.gz" > "$here/kc/$name.k$k.hist"
	done
done
# yak-count: what the reference prints for the same read sets, without and with the Bloom filter
# (one file: the result does not depend on the filter; -b 12 is too small for a filter to come of
# it, the second pass still runs), and with two files that share reads (there the filter's
# false positives show: -b 30 has none, -b 19 many)
rm -rf "$here/yak"; mkdir -p "$here/yak"
yak() { # out-name args...
	local out=$1; shift
	"$ref/yak-count" -t 2 "$@" > "$here/yak/$out.hist" 2>/dev/null
}
tmp=$(mktemp -d)
for name in k21 exotic; do
	fq="$here/e2e_$name/reads.fq.gz"
	yak "$name.k31" -k 31 "$fq"
	yak "$name.k21" -k 21 "$fq"
	yak "$name.k21.b24" -k 21 -b 24 "$fq"
	yak "$name.k15.b19.H3" -k 15 -b 19 -H 3 "$fq"
	yak "$name.k27.b12" -k 27 -b 12 "$fq"
	zcat "$fq" | head -n 12000 > "$tmp/a.fq"; zcat "$fq" | tail -n 12000 > "$tmp/b.fq"
	yak "$name.k21.b30.two" -k 21 -b 30 "$tmp/a.fq" "$tmp/b.fq"
	yak "$name.k21.b19.two" -k 21 -b 19 "$tmp/a.fq" "$tmp/b.fq"
	yak "$name.k21.b0.two" -k 21 "$tmp/a.fq" "$tmp/b.fq"
done
rm -rf "$tmp"
ls -la "$here" "$here"/e2e_* "$here"/kc "$here"/yak
