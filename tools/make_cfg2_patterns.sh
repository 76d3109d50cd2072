#!/bin/bash
# Fixture for BASELINE config 2: the reference's snp-pattern-gen -k 21 run over a synthetic
# hg38-length genome (tools/synth.c `genome`, seed 38, the panel's ref alleles planted) and the
# NGSCheckMate GRCh38 panel shipped with the reference.  Needs /root/reference and ~8 GB RAM,
# ~3 min.  Output: tests/golden/cfg2_patterns.txt.gz (20 797 of 20 920 SNPs survive; 27 rows
# repeat an earlier row's k-mers, which is what exercises first-insert-wins).
set -euo pipefail
root=$(cd "$(dirname "$0")/.." && pwd)
work=${1:-/tmp/cfg2}
bed=/root/reference/SNP/SNP_GRCh38_hg38_wChr.bed
make -s -C "$root/oracle" synth ref
mkdir -p "$work"
"$root/oracle/synth" genome -o "$work/hg38s.fa" -b "$bed" -s 38
"$root/oracle/_ref/snp-pattern-gen" -k 21 -b "$bed" -f "$work/hg38s.fa" -o "$work/cfg2_patterns.txt"
gzip -9nc "$work/cfg2_patterns.txt" > "$root/tests/golden/cfg2_patterns.txt.gz"
wc -l "$work/cfg2_patterns.txt"
