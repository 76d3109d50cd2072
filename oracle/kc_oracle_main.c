/*
 * kc_oracle_main.c -- command line over kc_oracle.c: `kc_oracle [-k INT] <in.fa>` prints the
 * histogram kc-c4 prints (kc-c4.c:217-252; -p / -t are accepted and ignored, the result does not
 * depend on them; -b only matters for a file with a malformed FASTQ record).  TEST INFRASTRUCTURE ONLY.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kc_oracle.h"

int main(int argc, char **argv)
{
	int k = 31, i;
	long block_len = 10000000;
	const char *fn = NULL;
	for (i = 1; i < argc; ++i) {
		if (argv[i][0] == '-' && argv[i][1] && strchr("kpbt", argv[i][1])) {
			const char opt = argv[i][1];
			const char *val = argv[i][2] ? argv[i] + 2 : (i + 1 < argc ? argv[++i] : "");
			if (opt == 'k') k = atoi(val);
			if (opt == 'b') block_len = atol(val);
		} else if (!fn) fn = argv[i];
	}
	if (!fn) {
		fprintf(stderr, "Usage: kc_oracle [-k INT] <in.fa>\n");
		return 1;
	}
	kco_t *o = kco_create(k);
	uint64_t hist[256];
	if (!o) return 1;
	kco_add_file(o, fn, block_len);
	kco_hist(o, hist);
	kco_print_hist(hist, stdout);
	fprintf(stderr, "[kc_oracle] %lu k-mers, %lu distinct\n", (unsigned long)kco_instances(o), (unsigned long)kco_distinct(o));
	kco_destroy(o);
	return 0;
}
