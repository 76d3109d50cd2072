/*
 * synth.c -- seeded synthetic inputs for vaf-counter: a reference FASTA, a SNP BED
 * and a FASTQ of reads drawn from a diploid donor that carries the SNPs.
 *
 * Everything is a pure function of (seed, index) through one 64-bit mixer, so the same
 * genome can be regenerated elsewhere (bench.py rebuilds it on the GPU with integer
 * tensor ops) without shipping files.
 *
 *   synth genome -o out.fa -b panel.bed [-s 38]
 *     an hg38-length random FASTA (3.1 Gb) with the BED's ref alleles planted
 *   synth cfg -o PREFIX [-L 1000000] [-n 1000] [-r 1000000] [-l 150] [-e 0.01]
 *             [-N 0.005] [-M 1] [-s 1] [-c chr1] [-j 0] [-x 0]
 *     writes PREFIX.fa, PREFIX.bed, PREFIX.fq
 *       -L genome length          -n number of SNPs         -r number of reads
 *       -l read length            -e substitution rate      -N rate at which an N run starts
 *       -M mean N-run length (geometric; 1 = single Ns)      -s seed
 *       -j jitter: read lengths uniform in [l-j, l+j]
 *       -x exotic rate: per base probability of an IUPAC / lower-case / odd byte
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

static inline uint64_t mix64(uint64_t z)
{
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

/* stateless stream: value i of stream `seed` */
static inline uint64_t at(uint64_t seed, uint64_t i)
{
	return mix64(seed + (i + 1) * 0x9E3779B97F4A7C15ULL);
}

typedef struct {
	uint64_t s, i;
} rng_t;

static inline uint64_t rnd(rng_t *r) { return at(r->s, r->i++); }
static inline double rnd01(rng_t *r) { return (double)(rnd(r) >> 11) * (1.0 / 9007199254740992.0); }

static const char BASES[4] = {'A', 'C', 'G', 'T'};

static inline char comp(char c)
{
	switch (c) {
	case 'A': return 'T';
	case 'C': return 'G';
	case 'G': return 'C';
	case 'T': return 'A';
	case 'a': return 't';
	case 'c': return 'g';
	case 'g': return 'c';
	case 't': return 'a';
	}
	return c;
}

/* hg38 primary contigs that carry SNPs of the NGSCheckMate panel, plus its two alt contigs */
static const struct { const char *name; long len; } HG38[] = {
	{"chr1", 248956422}, {"chr2", 242193529}, {"chr3", 198295559}, {"chr4", 190214555},
	{"chr5", 181538259}, {"chr6", 170805979}, {"chr7", 159345973}, {"chr8", 145138636},
	{"chr9", 138394717}, {"chr10", 133797422}, {"chr11", 135086622}, {"chr12", 133275309},
	{"chr13", 114364328}, {"chr14", 107043718}, {"chr15", 101991189}, {"chr16", 90338345},
	{"chr17", 83257441}, {"chr18", 80373285}, {"chr19", 58617616}, {"chr20", 64444167},
	{"chr21", 46709983}, {"chr22", 50818468}, {"chrX", 156040895},
	{"chr19_KI270938v1_alt", 1066800}, {"chr1_KI270766v1_alt", 256271},
};

/* synth genome -o out.fa -b panel.bed [-s seed]: an hg38-length random reference in which the
 * `ref` allele of every BED row is planted at its position (otherwise a quarter of the panel
 * would name an alt allele equal to the random base and be dropped by snp-pattern-gen). */
static int genome_main(int argc, char **argv)
{
	const char *out = NULL, *bed = NULL;
	uint64_t seed = 38;
	int c;
	optind = 2;
	while ((c = getopt(argc, argv, "o:b:s:")) >= 0) {
		if (c == 'o') out = optarg;
		else if (c == 'b') bed = optarg;
		else if (c == 's') seed = strtoull(optarg, 0, 10);
	}
	if (!out) { fprintf(stderr, "Usage: synth genome -o out.fa [-b panel.bed] [-s seed]\n"); return 1; }
	typedef struct { char chr[64]; long pos; char ref; } row_t;
	row_t *rows = NULL;
	long n_rows = 0, m_rows = 0;
	if (bed) {
		FILE *fb = fopen(bed, "r");
		char chr[256], id[256], r, a;
		long st, en;
		if (!fb) { perror(bed); return 1; }
		while (fscanf(fb, "%255s%ld%ld%255s %c %c", chr, &st, &en, id, &r, &a) == 6) {
			if (n_rows == m_rows) rows = (row_t *)realloc(rows, (m_rows = m_rows ? m_rows * 2 : 1024) * sizeof(row_t));
			snprintf(rows[n_rows].chr, sizeof rows[n_rows].chr, "%.63s", chr);
			rows[n_rows].pos = st;
			rows[n_rows++].ref = r;
		}
		fclose(fb);
	}
	FILE *fp = fopen(out, "w");
	if (!fp) { perror(out); return 1; }
	static char obuf[1 << 22];
	setvbuf(fp, obuf, _IOFBF, sizeof obuf);
	uint64_t base = 0;
	long planted = 0;
	for (unsigned k = 0; k < sizeof HG38 / sizeof *HG38; ++k) {
		long L = HG38[k].len;
		char *s = (char *)malloc((size_t)L);
		for (long i = 0; i < L; ++i) s[i] = BASES[at(seed, base + (uint64_t)i) >> 62];
		for (long j = 0; j < n_rows; ++j)
			if (strcmp(rows[j].chr, HG38[k].name) == 0 && rows[j].pos >= 0 && rows[j].pos < L &&
			    strchr("ACGT", rows[j].ref)) {
				s[rows[j].pos] = rows[j].ref;
				++planted;
			}
		fprintf(fp, ">%s\n", HG38[k].name);
		for (long i = 0; i < L; i += 60) {
			fwrite(s + i, 1, (size_t)(L - i < 60 ? L - i : 60), fp);
			fputc('\n', fp);
		}
		free(s);
		base += (uint64_t)L;
	}
	fclose(fp);
	fprintf(stderr, "synth genome: %llu bases, %ld alleles planted\n", (unsigned long long)base, planted);
	free(rows);
	return 0;
}

int main(int argc, char **argv)
{
	long L = 1000000, n_snp = 1000, n_reads = 1000000;
	int rl = 150, jitter = 0, c;
	double e_sub = 0.01, e_n = 0.005, n_mean = 1.0, e_exotic = 0.0;
	uint64_t seed = 1;
	const char *prefix = NULL, *chr = "chr1";
	char path[4096];
	static const char EXOTIC[] = "RYKMSWBDHVacgtnUu.-*XQ17";

	if (argc >= 2 && strcmp(argv[1], "genome") == 0) return genome_main(argc, argv);
	if (argc < 2 || strcmp(argv[1], "cfg") != 0) {
		fprintf(stderr, "Usage: synth cfg -o PREFIX [-L len] [-n snps] [-r reads] [-l readlen] [-e sub] [-N nrate] [-M nrun] [-s seed] [-c chr] [-j jitter] [-x exotic]\n");
		return 1;
	}
	optind = 2;
	while ((c = getopt(argc, argv, "o:L:n:r:l:e:N:M:s:c:j:x:")) >= 0) {
		if (c == 'o') prefix = optarg;
		else if (c == 'L') L = atol(optarg);
		else if (c == 'n') n_snp = atol(optarg);
		else if (c == 'r') n_reads = atol(optarg);
		else if (c == 'l') rl = atoi(optarg);
		else if (c == 'e') e_sub = atof(optarg);
		else if (c == 'N') e_n = atof(optarg);
		else if (c == 'M') n_mean = atof(optarg);
		else if (c == 's') seed = strtoull(optarg, 0, 10);
		else if (c == 'c') chr = optarg;
		else if (c == 'j') jitter = atoi(optarg);
		else if (c == 'x') e_exotic = atof(optarg);
	}
	if (!prefix || L < 1000 || rl + jitter > L || rl - jitter < 1) {
		fprintf(stderr, "synth: bad arguments\n");
		return 1;
	}

	/* reference: base i = top two bits of stream (seed) at i */
	char *ref = (char *)malloc((size_t)L + 1);
	for (long i = 0; i < L; ++i) ref[i] = BASES[at(seed, (uint64_t)i) >> 62];
	ref[L] = 0;
	snprintf(path, sizeof path, "%s.fa", prefix);
	FILE *fp = fopen(path, "w");
	if (!fp) { perror(path); return 1; }
	fprintf(fp, ">%s synthetic seed=%llu\n", chr, (unsigned long long)seed);
	for (long i = 0; i < L; i += 60) fprintf(fp, "%.*s\n", (int)(L - i < 60 ? L - i : 60), ref + i);
	fclose(fp);

	/* SNPs: distinct positions in [100, L-100), alt = one of the other three bases */
	char *h0 = (char *)malloc((size_t)L), *h1 = (char *)malloc((size_t)L);
	memcpy(h0, ref, (size_t)L);
	memcpy(h1, ref, (size_t)L);
	uint8_t *taken = (uint8_t *)calloc((size_t)L, 1);
	rng_t rs = {mix64(seed ^ 0x534e50ULL), 0};
	snprintf(path, sizeof path, "%s.bed", prefix);
	fp = fopen(path, "w");
	if (!fp) { perror(path); return 1; }
	for (long j = 0; j < n_snp; ++j) {
		long pos;
		do pos = 100 + (long)(rnd(&rs) % (uint64_t)(L - 200)); while (taken[pos]);
		taken[pos] = 1;
		char r = ref[pos], a;
		do a = BASES[rnd(&rs) & 3]; while (a == r);
		int g = (int)(rnd(&rs) & 3); /* 0: hom-ref, 1,2: het, 3: hom-alt */
		if (g >= 1) h1[pos] = a;
		if (g == 3) h0[pos] = a;
		fprintf(fp, "%s\t%ld\t%ld\trs%ld\t%c\t%c\n", chr, pos, pos + 1, j, r, a);
	}
	fclose(fp);
	free(taken);

	/* reads */
	snprintf(path, sizeof path, "%s.fq", prefix);
	fp = fopen(path, "w");
	if (!fp) { perror(path); return 1; }
	static char obuf[1 << 22];
	setvbuf(fp, obuf, _IOFBF, sizeof obuf);
	int maxl = rl + jitter;
	char *s = (char *)malloc((size_t)maxl + 1), *q = (char *)malloc((size_t)maxl + 1);
	memset(q, 'I', (size_t)maxl);
	double p_end = n_mean > 1.0 ? 1.0 / n_mean : 1.0;
	for (long j = 0; j < n_reads; ++j) {
		rng_t r = {mix64(seed ^ 0x52454144ULL) + (uint64_t)j * 0x632BE59BD9B4E019ULL, 0};
		int len = jitter ? rl - jitter + (int)(rnd(&r) % (uint64_t)(2 * jitter + 1)) : rl;
		long st = (long)(rnd(&r) % (uint64_t)(L - len + 1));
		uint64_t fl = rnd(&r);
		const char *h = (fl & 1) ? h1 : h0;
		int rev = (int)(fl >> 1 & 1), in_n = 0;
		for (int i = 0; i < len; ++i) {
			char b = h[st + i];
			if (e_sub > 0 && rnd01(&r) < e_sub) {
				char nb;
				do nb = BASES[rnd(&r) & 3]; while (nb == b);
				b = nb;
			}
			if (in_n) {
				b = 'N';
				if (rnd01(&r) < p_end) in_n = 0;
			} else if (e_n > 0 && rnd01(&r) < e_n) {
				b = 'N';
				if (n_mean > 1.0 && rnd01(&r) >= p_end) in_n = 1;
			}
			if (e_exotic > 0 && rnd01(&r) < e_exotic) b = EXOTIC[rnd(&r) % (sizeof(EXOTIC) - 1)];
			s[i] = b;
		}
		if (rev) {
			for (int i = 0, k2 = len - 1; i < k2; ++i, --k2) {
				char t = comp(s[i]);
				s[i] = comp(s[k2]);
				s[k2] = t;
			}
			if (len & 1) s[len / 2] = comp(s[len / 2]);
		}
		fprintf(fp, "@r%ld\n%.*s\n+\n%.*s\n", j, len, s, len, q);
	}
	fclose(fp);
	free(s); free(q); free(ref); free(h0); free(h1);
	return 0;
}
