/*
 * ref_kat.c -- known-answer dumper.  TEST INFRASTRUCTURE ONLY.
 *
 * Compiles the reference's vaf-counter translation unit INTO this program by path
 * (-DREF_TU='"/root/reference/vaf-counter.c"', main renamed) so its file-static functions
 * -- kmer_hash, encode_seq_simd, extract_kmers_to_buf, the khashl instance -- can be
 * called unmodified.  Output (TSV on stdout) is committed as tests/golden/kat_vaf.tsv by
 * tests/golden/make_golden.sh; the oracle and the CUDA path are checked against it.
 */
#define main vc_reference_main
#include REF_TU
#undef main

static uint64_t s_rng = 0x243F6A8885A308D3ULL;
static uint64_t rnd(void)
{
	uint64_t z = (s_rng += 0x9E3779B97F4A7C15ULL);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

static void dump_kmer(const char *s, int k)
{
	uint64_t f = encode_kmer(s, k);
	if (f == UINT64_MAX) {
		printf("kmer\t%d\t%.*s\tinvalid\n", k, k, s);
		return;
	}
	uint64_t r = revcomp_kmer(f, k), c = canonical_kmer(f, k);
	printf("kmer\t%d\t%.*s\t%llx\t%llx\t%llx\t%x\n", k, k, s, (unsigned long long)f,
	       (unsigned long long)r, (unsigned long long)c, kmer_hash(c));
}

static void dump_extract(int k, const char *seq, int len)
{
	kmer_buf_t b = {0, 0, 0};
	extract_kmers_to_buf(&b, k, len, seq);
	printf("extract\t%d\t%d\t", k, len);
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
	char buf[512];
	int ks[] = {1, 2, 3, 11, 15, 16, 21, 27, 31};

	/* 1. byte classes: strict table and the 16-byte SIMD encoder */
	for (int b = 0; b < 256; ++b) {
		char in[16];
		uint8_t out[16];
		memset(in, b, 16);
		encode_seq_simd(in, 16, out);
		printf("nt4\t%d\t%d\t%d\n", b, seq_nt4_table[b], out[0]);
	}
	/* 2. k-mer arithmetic */
	dump_kmer("CACTCAAACACTCGGACCGGC", 21);
	dump_kmer("ACGTACGTACGTACGTACGTA", 21);
	dump_kmer("TTTTTTTTTTTTTTTTTTTTT", 21);
	dump_kmer("AAAAAAAAAAAAAAAAAAAAA", 21);
	dump_kmer("ACGTTGCAAGGCTTAACCGGTTACGATCGAT", 31);
	dump_kmer("acgtugcaaggcuuaaccgguuacgaucgau", 31);
	dump_kmer("ACGTNGCAAGGCTTAACCGGT", 21);
	for (unsigned i = 0; i < sizeof ks / sizeof *ks; ++i)
		for (int j = 0; j < 24; ++j) {
			for (int t = 0; t < ks[i]; ++t) buf[t] = ACGT[rnd() & 3];
			dump_kmer(buf, ks[i]);
		}
	/* 3. bucket function */
	for (int j = 0; j < 64; ++j) {
		khint_t h = (khint_t)rnd();
		khint_t bits = 2 + (khint_t)(rnd() % 24);
		printf("h2b\t%x\t%u\t%u\n", h, bits, __kh_h2b(h, bits));
	}
	/* 4. extractor on reads of every length class, with Ns and odd bytes */
	for (unsigned i = 0; i < sizeof ks / sizeof *ks; ++i) {
		int k = ks[i];
		for (int j = 0; j < 40; ++j) {
			int len = (int)(rnd() % 100);
			int mode = j % 4; /* 0: pure ACGT, 1: 3% N, 2: 10% mixed junk, 3: N runs */
			for (int t = 0; t < len; ++t) {
				char c = ACGT[rnd() & 3];
				if (mode == 1 && rnd() % 100 < 3) c = 'N';
				if (mode == 2 && rnd() % 100 < 10) c = MIXED[rnd() % (sizeof(MIXED) - 1)];
				if (mode == 3 && t > 0 && buf[t - 1] == 'N' && rnd() % 100 < 70) c = 'N';
				else if (mode == 3 && rnd() % 100 < 4) c = 'N';
				buf[t] = c;
			}
			dump_extract(k, buf, len);
		}
		/* lengths around the 16-byte SIMD boundary, junk everywhere */
		for (int len = k > 2 ? k - 2 : 0; len <= k + 36; ++len) {
			for (int t = 0; t < len; ++t)
				buf[t] = rnd() % 100 < 12 ? MIXED[rnd() % (sizeof(MIXED) - 1)] : ACGT[rnd() & 3];
			dump_extract(k, buf, len);
		}
	}
	/* 5. table geometry and first-insert-wins (vaf-counter.c:198-252) */
	int ns[] = {0, 1, 2, 5, 341, 342, 1000, 1365, 1366, 20920};
	for (unsigned i = 0; i < sizeof ns / sizeof *ns; ++i) {
		pattern_db_t db = {0, 0, 0};
		db.n = db.m = ns[i];
		db.a = (pattern_t *)calloc(db.n ? db.n : 1, sizeof(pattern_t));
		for (int p = 0; p < db.n; ++p) {
			for (int t = 0; t < 21; ++t) db.a[p].ref_kmer[t] = ACGT[rnd() & 3];
			memcpy(db.a[p].alt_kmer, db.a[p].ref_kmer, 21);
			db.a[p].alt_kmer[10] = ACGT[(strchr(ACGT, db.a[p].ref_kmer[10]) - ACGT + 1 + rnd() % 3) & 3];
			if (p % 97 == 5 && p > 0) memcpy(db.a[p].ref_kmer, db.a[p - 1].ref_kmer, 21); /* duplicate */
			if (p % 131 == 7) db.a[p].alt_kmer[3] = 'N';                                  /* unusable */
		}
		kmer_cnt_t *h = create_combined_kmer_map(&db, 21);
		printf("map\t%d\t%u\t%u", db.n, h->bits, kh_size(h));
		/* probe a few keys: value or -1 */
		for (int p = 0; p < db.n && p < 12; ++p) {
			uint64_t x = encode_kmer(db.a[p].ref_kmer, 21);
			khint_t it = kmer_cnt_get(h, canonical_kmer(x, 21));
			printf("\t%s:%d", db.a[p].ref_kmer, it == kh_end(h) ? -1 : (int)kh_val(h, it));
		}
		printf("\n");
		kmer_cnt_destroy(h);
		free(db.a);
	}
	return 0;
}
