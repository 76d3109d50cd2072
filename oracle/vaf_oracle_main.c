/*
 * vaf_oracle_main.c -- command-line front end of the CPU restatement.
 * TEST INFRASTRUCTURE ONLY (see vaf_oracle.h).  Same flags as the reference tool
 * (vaf-counter.c:598-617) plus -S 0|1 to pick the scalar / SSSE3 encoder rule.
 */
#include <stdio.h>
#include <stdlib.h>
#include <sys/time.h>
#include <unistd.h>

#include "vaf_oracle.h"

static double now(void)
{
	struct timeval tv;
	gettimeofday(&tv, NULL);
	return tv.tv_sec + tv.tv_usec * 1e-6;
}

int main(int argc, char **argv)
{
	int c, k = 21, nt = 4, block = 10000000, simd = 1, verbose = 0;
	const char *pfn = NULL, *ofn = NULL;
	while ((c = getopt(argc, argv, "k:p:o:t:b:vS:")) >= 0) {
		if (c == 'k') k = atoi(optarg);
		else if (c == 'p') pfn = optarg;
		else if (c == 'o') ofn = optarg;
		else if (c == 't') nt = atoi(optarg);
		else if (c == 'b') block = atoi(optarg);
		else if (c == 'v') verbose = 1;
		else if (c == 'S') simd = atoi(optarg);
	}
	if (!pfn || !ofn || optind >= argc || k < 1 || k > 31) {
		fprintf(stderr, "Usage: vaf_oracle [-k 21] [-t 4] [-b 10000000] [-S 1] [-v] -p patterns.txt -o out.vaf reads.fq [...]\n");
		return 1;
	}
	vo_patterns_t *db = vo_load_patterns(pfn);
	if (!db) {
		fprintf(stderr, "Error: failed to load pattern file\n");
		return 1;
	}
	vo_map_t *m = vo_map_build(db, k);
	uint32_t *counts = (uint32_t *)calloc((size_t)2 * (db->n > 0 ? db->n : 1), 4);
	vo_stats_t st = {0, 0, 0};
	double t0 = now();
	for (int i = optind; i < argc; ++i)
		vo_count_file(m, k, argv[i], simd, nt, block, counts, &st);
	double dt = now() - t0;
	FILE *fp = fopen(ofn, "w");
	if (!fp) {
		fprintf(stderr, "Error: failed to open output file\n");
		return 1;
	}
	vo_write_vaf(fp, db, counts);
	fclose(fp);
	if (verbose)
		fprintf(stderr, "[oracle] %llu reads, %llu bases, %llu k-mers, %d collisions, %.3f s, %.2f Mbases/s (%d threads)\n",
		        (unsigned long long)st.n_reads, (unsigned long long)st.n_bases,
		        (unsigned long long)st.n_kmers, m->n_collisions, dt, st.n_bases / dt / 1e6, nt);
	free(counts);
	vo_map_free(m);
	vo_patterns_free(db);
	return 0;
}
