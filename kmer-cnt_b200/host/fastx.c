/*
 * fastx.c -- see fastx.h.  Record rules follow kseq.h:192-232 of the reference; the
 * implementation (one large refill buffer, memchr line scanning) is our own.
 */
#include "fastx.h"

#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#define FX_BUFSZ (1u << 20)

typedef struct {
	char *s;
	size_t l, m;
} fx_str_t;

struct fastx {
	gzFile fp;
	unsigned char *buf;
	size_t beg, end;
	int eof;
	int header_seen; /* a '>' or '@' has been consumed and its name follows */
	fx_str_t seq, name;
	size_t qual_len;
	int qual_last; /* last quality byte gathered so far, -1 if none */
};

static int fx_fill(fastx_t *fx)
{
	int n;
	if (fx->eof) return 0;
	n = gzread(fx->fp, fx->buf, FX_BUFSZ);
	fx->beg = 0;
	if (n <= 0) { /* a read error ends the stream, as a short read does in kseq.h:85 */
		fx->end = 0;
		fx->eof = 1;
		return 0;
	}
	fx->end = (size_t)n;
	return 1;
}

static inline int fx_getc(fastx_t *fx)
{
	if (fx->beg >= fx->end && !fx_fill(fx)) return -1;
	return fx->buf[fx->beg++];
}

static void fx_reserve(fx_str_t *s, size_t extra)
{
	if (s->l + extra + 1 > s->m) {
		size_t m = s->m ? s->m : 256;
		while (m < s->l + extra + 1) m <<= 1;
		s->s = (char *)realloc(s->s, m);
		s->m = m;
	}
}

/* Consume up to and including the next '\n'.  If dst != NULL the line is appended to
 * it, else only its length is added to *count and its last byte kept in *last.  After
 * each line a '\r' at the very end of what has been gathered so far is dropped when
 * more than one byte has been gathered (kseq.h:146).  Returns 0 if the stream was
 * already exhausted on entry, 1 otherwise. */
static int fx_line(fastx_t *fx, fx_str_t *dst, size_t *count, int *last)
{
	if (fx->beg >= fx->end && !fx_fill(fx)) return 0;
	for (;;) {
		unsigned char *p = fx->buf + fx->beg;
		size_t avail = fx->end - fx->beg;
		unsigned char *nl = (unsigned char *)memchr(p, '\n', avail);
		size_t n = nl ? (size_t)(nl - p) : avail;
		if (n) {
			if (dst) {
				fx_reserve(dst, n);
				memcpy(dst->s + dst->l, p, n);
				dst->l += n;
			} else {
				*count += n;
				*last = p[n - 1];
			}
		}
		fx->beg += n + (nl ? 1 : 0);
		if (nl) break;
		if (!fx_fill(fx)) break;
	}
	if (dst) {
		if (dst->l > 1 && dst->s[dst->l - 1] == '\r') --dst->l;
	} else if (*count > 1 && *last == '\r') {
		--*count;
		*last = -1; /* whatever precedes a dropped '\r' is never looked at again */
	}
	return 1;
}

fastx_t *fastx_open(const char *fn)
{
	fastx_t *fx;
	gzFile fp = gzopen(fn, "r");
	if (!fp) return NULL;
	gzbuffer(fp, 1u << 20);
	fx = (fastx_t *)calloc(1, sizeof(*fx));
	fx->fp = fp;
	fx->buf = (unsigned char *)malloc(FX_BUFSZ);
	return fx;
}

void fastx_close(fastx_t *fx)
{
	if (!fx) return;
	gzclose(fx->fp);
	free(fx->buf);
	free(fx->seq.s);
	free(fx->name.s);
	free(fx);
}

static inline int fx_isspace(int c)
{
	return c == ' ' || (c >= '\t' && c <= '\r');
}

const char *fastx_name(const fastx_t *fx) { return fx->name.s ? fx->name.s : ""; }

long fastx_next(fastx_t *fx, const char **seq)
{
	int c;
	if (!fx->header_seen) { /* hunt for the next header character, anywhere */
		while ((c = fx_getc(fx)) != -1 && c != '>' && c != '@') {}
		if (c == -1) return -1;
		fx->header_seen = 1;
	}
	fx->seq.l = 0;
	fx->qual_len = 0;
	fx->qual_last = -1;
	/* name: up to the first white space; nothing left at all means end of input */
	if (fx->beg >= fx->end && !fx_fill(fx)) return -1;
	fx->name.l = 0;
	while ((c = fx_getc(fx)) != -1 && !fx_isspace(c)) {
		fx_reserve(&fx->name, 1);
		fx->name.s[fx->name.l++] = (char)c;
	}
	fx_reserve(&fx->name, 0);
	fx->name.s[fx->name.l] = 0;
	if (c != '\n' && c != -1) { /* comment: rest of the header line */
		size_t dummy = 0;
		int dlast = -1;
		fx_line(fx, NULL, &dummy, &dlast);
	}
	/* sequence lines until a line starts with '>', '@' or '+' */
	while ((c = fx_getc(fx)) != -1 && c != '>' && c != '+' && c != '@') {
		if (c == '\n') continue;
		fx_reserve(&fx->seq, 1);
		fx->seq.s[fx->seq.l++] = (char)c;
		fx_line(fx, &fx->seq, NULL, NULL);
	}
	/* header_seen stays set: either the next header character has just been consumed,
	 * or the input ended (the next call then reports -1), or a '+' line follows */
	*seq = fx->seq.s ? fx->seq.s : "";
	if (c != '+') return (long)fx->seq.l; /* FASTA record (or end of input) */
	/* FASTQ: drop the rest of the '+' line, then gather quality up to the sequence length */
	while ((c = fx_getc(fx)) != -1 && c != '\n') {}
	if (c == -1) return -2; /* header_seen stays set: the next call returns -1 */
	while (fx_line(fx, NULL, &fx->qual_len, &fx->qual_last) && fx->qual_len < fx->seq.l) {}
	fx->header_seen = 0;
	if (fx->qual_len != fx->seq.l) return -2;
	return (long)fx->seq.l;
}
