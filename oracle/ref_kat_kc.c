/*
 * ref_kat_kc.c -- known-answer dumper for the counting mode.  TEST INFRASTRUCTURE ONLY.
 *
 * Compiles the reference's kc-c4 translation unit INTO this program by path
 * (-DREF_TU='"/root/reference/kc-c4.c"', main renamed) so that its file-static hash64 and
 * count_seq_buf can be called unmodified.  Output (TSV on stdout) is committed as
 * tests/golden/kat_kc.tsv by tests/golden/make_golden.sh; the oracle and the CUDA path are
 * checked against it.
 */
#define main kc_reference_main
#include REF_TU
#undef main

static uint64_t s_rng = 0x13198A2E03707344ULL;
static uint64_t rnd(void)
{
	uint64_t z = (s_rng += 0x9E3779B97F4A7C15ULL);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

/* count_seq_buf with p = 0 files every hashed k-mer of the read, in order, in one buffer */
static void dump_read(int k, const char *seq, int len)
{
	buf_c4_t b = {0, 0, 0};
	count_seq_buf(&b, k, 0, len, seq);
	printf("kmers\t%d\t", k);
	for (int i = 0; i < len; ++i) printf("%02x", (unsigned char)seq[i]);
	printf("\t%d\t", b.n);
	for (int i = 0; i < b.n; ++i) printf("%s%llx", i ? "," : "", (unsigned long long)b.a[i]);
	printf("\n");
	free(b.a);
}

int main(void)
{
	static const char ACGT[] = "ACGT";
	static const char MIXED[] = "ACGTACGTACGTACGTNNacgtnUuRYKMSWBDHV.-*Q1357 \t\x01\x02\x03";
	int ks[] = {1, 2, 5, 11, 15, 16, 21, 27, 28, 31};
	char buf[400];
	for (int b = 0; b < 256; ++b) printf("nt4\t%d\t%d\n", b, seq_nt4_table[b]);
	for (unsigned i = 0; i < sizeof ks / sizeof *ks; ++i) {
		const int k = ks[i];
		const uint64_t mask = (1ULL << k * 2) - 1;
		uint64_t edge[] = {0, 1, mask, mask - 1, mask >> 1, 0x5555555555555555ULL & mask};
		for (unsigned j = 0; j < sizeof edge / sizeof *edge; ++j)
			printf("hash64\t%d\t%llx\t%llx\n", k, (unsigned long long)edge[j], (unsigned long long)hash64(edge[j], mask));
		for (int j = 0; j < 40; ++j) {
			uint64_t x = rnd() & mask;
			printf("hash64\t%d\t%llx\t%llx\n", k, (unsigned long long)x, (unsigned long long)hash64(x, mask));
		}
		for (int j = 0; j < 6; ++j) {
			int len = j == 0 ? k : j == 1 ? k - 1 : (int)(k + rnd() % 150);
			for (int t = 0; t < len; ++t) {
				uint64_t r = rnd();
				buf[t] = (r >> 8) % 100 < 4 ? MIXED[(r >> 20) % (sizeof MIXED - 1)] : ACGT[r & 3];
			}
			dump_read(k, buf, len);
		}
	}
	return 0;
}
