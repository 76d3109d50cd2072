/*
 * yak_oracle_main.c -- command line over yak_oracle.c: `yak_oracle [-k -p -b -H -K -t] <in.fa> [in2.fa]`
 * prints what yak-count prints (yak-count.c:460-507; -t is accepted and ignored, the result
 * does not depend on it).  TEST INFRASTRUCTURE ONLY.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "yak_oracle.h"

int main(int argc, char **argv)
{
	int k = 31, pre = 10, bf_shift = 0, n_hash = 4, i;
	long chunk = 10000000;
	const char *fn[2] = {NULL, NULL};
	for (i = 1; i < argc; ++i) {
		if (argv[i][0] == '-' && argv[i][1] && strchr("kpKtbH", argv[i][1])) {
			const char opt = argv[i][1];
			const char *val = argv[i][2] ? argv[i] + 2 : (i + 1 < argc ? argv[++i] : "");
			if (opt == 'k') k = atoi(val);
			if (opt == 'p') pre = atoi(val);
			if (opt == 'K') chunk = atoi(val);
			if (opt == 'b') bf_shift = atoi(val);
			if (opt == 'H') n_hash = atoi(val);
		} else if (!fn[0]) fn[0] = argv[i];
		else if (!fn[1]) fn[1] = argv[i];
	}
	if (!fn[0]) {
		fprintf(stderr, "Usage: yak_oracle [options] <in.fa> [in.fa]\n");
		return 1;
	}
	if (pre < 10) {
		fprintf(stderr, "ERROR: -p should be at least %d\n", 10);
		return 1;
	}
	yko_t *o = yko_create(k, pre, bf_shift, n_hash);
	uint64_t hist[1024];
	if (!o || yko_count_files(o, fn[0], fn[1], chunk) != 0) return 1;
	fprintf(stderr, "[yak_oracle] %ld distinct k-mers after shrinking\n", (long)yko_distinct(o));
	yko_hist(o, hist);
	yko_print_hist(hist, stdout);
	yko_destroy(o);
	return 0;
}
