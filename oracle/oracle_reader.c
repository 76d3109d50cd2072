/* oracle_reader.c -- see oracle_reader.h.  TEST INFRASTRUCTURE ONLY. */
#include "oracle_reader.h"

#include <ctype.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

struct orr {
	char *text;      /* the whole input, decompressed */
	size_t n, at;    /* its length, the read position */
	int pending;     /* header character already consumed by the previous record (kseq's last_char) */
	char *seq, *qual, *name;
	size_t seq_l, seq_m, qual_l, qual_m, name_m;
};

orr_t *orr_open(const char *fn)
{
	gzFile fp = gzopen(fn, "r");
	if (!fp) return NULL;
	orr_t *r = calloc(1, sizeof *r);
	size_t cap = 1 << 20;
	r->text = malloc(cap);
	for (;;) {
		if (r->n == cap) r->text = realloc(r->text, cap *= 2);
		int got = gzread(fp, r->text + r->n, (unsigned)(cap - r->n > (1u << 30) ? (1u << 30) : cap - r->n));
		if (got <= 0) break;
		r->n += (size_t)got;
	}
	gzclose(fp);
	r->name = calloc(1, r->name_m = 64);
	return r;
}

void orr_close(orr_t *r)
{
	if (!r) return;
	free(r->text), free(r->seq), free(r->qual), free(r->name), free(r);
}

static int next_char(orr_t *r) { return r->at < r->n ? (unsigned char)r->text[r->at++] : -1; }

static void grow(char **s, size_t *m, size_t need)
{
	if (need <= *m) return;
	while (*m < need) *m = *m ? *m * 2 : 256;
	*s = realloc(*s, *m);
}

/* ks_getuntil2(..., KS_SEP_LINE, ..., append = 1): the rest of the current line onto (*s, *l);
 * a trailing '\r' is dropped when the string is longer than one character (kseq.h:139).
 * Returns -1 when the input was already exhausted. */
static long append_line(orr_t *r, char **s, size_t *l, size_t *m)
{
	if (r->at >= r->n) return -1;
	const char *from = r->text + r->at;
	const char *nl = memchr(from, '\n', r->n - r->at);
	size_t len = nl ? (size_t)(nl - from) : r->n - r->at;
	grow(s, m, *l + len + 1);
	memcpy(*s + *l, from, len);
	*l += len;
	r->at += len + (nl ? 1 : 0);
	if (*l > 1 && (*s)[*l - 1] == '\r') --*l;
	return (long)*l;
}

long orr_next(orr_t *r, const char **seq)
{
	int c;
	if (!r->pending) { /* kseq.h:195-199: on to the next '>' or '@', wherever it stands */
		while ((c = next_char(r)) != -1 && c != '>' && c != '@') {}
		if (c == -1) return -1;
		r->pending = c;
	}
	r->seq_l = r->qual_l = 0;
	/* kseq.h:201-202: the name up to white space, the comment up to the end of the line */
	if (r->at >= r->n) return -1;
	size_t e = r->at;
	while (e < r->n && !isspace((unsigned char)r->text[e])) ++e;
	grow(&r->name, &r->name_m, e - r->at + 1);
	memcpy(r->name, r->text + r->at, e - r->at);
	r->name[e - r->at] = 0;
	c = e < r->n ? (unsigned char)r->text[e] : 0;
	r->at = e < r->n ? e + 1 : r->n;
	if (c != '\n') {
		const char *nl = r->at < r->n ? memchr(r->text + r->at, '\n', r->n - r->at) : NULL;
		r->at = nl ? (size_t)(nl - r->text) + 1 : r->n;
	}
	/* kseq.h:207-211: sequence lines until one starts with '>', '+' or '@'; empty lines skipped */
	grow(&r->seq, &r->seq_m, 256);
	while ((c = next_char(r)) != -1 && c != '>' && c != '+' && c != '@') {
		if (c == '\n') continue;
		grow(&r->seq, &r->seq_m, r->seq_l + 2);
		r->seq[r->seq_l++] = (char)c;
		append_line(r, &r->seq, &r->seq_l, &r->seq_m);
	}
	if (c == '>' || c == '@') r->pending = c;
	*seq = r->seq;
	if (c != '+') return (long)r->seq_l; /* FASTA (kseq.h:219) */
	/* kseq.h:224-230: skip the '+' line, then quality lines until they are as long as the sequence */
	while ((c = next_char(r)) != -1 && c != '\n') {}
	if (c == -1) return -2;
	while (append_line(r, &r->qual, &r->qual_l, &r->qual_m) >= 0 && r->qual_l < r->seq_l) {}
	r->pending = 0;
	if (r->seq_l != r->qual_l) return -2;
	return (long)r->seq_l;
}

const char *orr_name(const orr_t *r) { return r->name; }
