/*
 * anchor_sim.cpp -- TEST-ONLY host emulation of anchor_scan_kernel (vafgpu_kernels.cu).
 *
 * Lets the CPU test-suite check the table builder (vafgpu_tables.cpp), the packing
 * arithmetic and the anchor bookkeeping against the oracle without a GPU.  It is built by
 * tests/conftest.py into tests/_build/ and is never linked into libvafgpu.so: the product
 * has no CPU path.
 */
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../kmer-cnt_b200/csrc/vafgpu_tables.hpp"

using namespace vafgpu;

static inline bool is_base(uint32_t b)
{
	uint32_t u = b & 0xDFu;
	return u == 'A' || u == 'C' || u == 'G' || u == 'T' || u == 'U';
}

static inline uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) /* __byte_perm */
{
	uint64_t v = (uint64_t)b << 32 | a;
	uint32_t r = 0;
	for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> 8 * ((sel >> 4 * i) & 7)) & 0xFF) << 8 * i;
	return r;
}

static inline uint32_t pack16(const uint8_t *c)
{
	const uint32_t M = 0x00820820u;
	uint32_t w[4], p[4];
	memcpy(w, c, 16);
	for (int i = 0; i < 4; ++i) p[i] = (w[i] & 0x06060606u) * M;
	uint32_t lo = prmt(p[0], p[1], 0x0073), hi = prmt(p[2], p[3], 0x0073);
	return prmt(lo, hi, 0x5410);
}

static inline uint32_t funnel_r(uint32_t lo, uint32_t hi, int sh)
{
	sh &= 31;
	return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}

extern "C" {

/* counts[val] += occurrences; returns filter survivors.  bytes: stream, n_bytes % 16 == 0 */
uint64_t sim_anchor_count(int k, const uint64_t *keys, const uint32_t *vals, uint32_t n,
                          const uint8_t *bytes, uint64_t n_bytes, uint32_t *counts,
                          uint32_t *info /* [stride, len, filter_words, slot_bits, entries, filter_keys] */)
{
	AnchorTables t;
	build_anchor_tables(k, keys, vals, n, t);
	const int S = t.plan.stride, L = t.plan.len;
	const uint32_t amask = vg_mask32(L);
	const uint32_t nw = t.filter_words;
	if (info) {
		info[0] = S, info[1] = L, info[2] = nw, info[3] = t.n_buckets;
		info[4] = t.n_entries, info[5] = t.n_filter_keys | (t.canon ? 0x80000000u : 0);
	}
	uint64_t n_cand = 0;
	const uint64_t n_chunks = n_bytes / 16;
	static const uint8_t NL[16] = {10, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10};
	for (uint64_t c = 0; c < n_chunks; ++c) {
		uint32_t cur = pack16(bytes + 16 * c);
		uint32_t left = pack16(c > 0 ? bytes + 16 * (c - 1) : NL);
		for (int j = 0; j < 16 / S; ++j) {
			const int s0 = (j + 1) * S - L; /* anchor = bases [s0, s0 + L) relative to the chunk */
			uint32_t a = (s0 >= 0 ? cur >> 2 * s0 : funnel_r(left, cur, 2 * (s0 + 16))) & amask;
			uint32_t key = vg_filter_key(a, L, t.canon);
			uint32_t m = vg_filter_mask(key, nw);
			if ((t.filter[vg_filter_word(key, nw)] & m) != m) continue;
			++n_cand;
			uint64_t q = 16 * c + (uint64_t)(j + 1) * S; /* aligned end of the anchor */
			for (uint32_t bk = vg_bucket_home(vg_hash_lo(key, nw), t.n_buckets);; bk = bk + 1 == t.n_buckets ? 0 : bk + 1) {
				const uint32_t ctrl = t.buckets[(size_t)bk * 4 + 3];
				for (int i = 0; i < 3; ++i) {
					if (t.buckets[(size_t)bk * 4 + i] != vg_tag(a, L)) continue;
					const vg_slot_t &e = t.slots[(ctrl & ~VG_CTRL_MORE) + i];
					const uint64_t end = q + e.off;
					if (((uint32_t)(e.okey >> 2 * (k - e.off - L)) & amask) != a || end < (uint64_t)k || end > n_bytes) continue;
					const uint8_t *b = bytes + (end - k);
					uint64_t km = 0;
					bool ok = true;
					for (int x = 0; x < k; ++x) {
						ok &= is_base(b[x]);
						km |= (uint64_t)((b[x] >> 1) & 3u) << 2 * x;
					}
					if (ok && km == e.okey) ++counts[e.val];
				}
				if (!(ctrl & VG_CTRL_MORE)) break;
			}
		}
	}
	return n_cand;
}

/* recipe-table membership, for checking build_recipe_table against the oracle's map */
int sim_recipe_get(int k, const uint64_t *keys, const uint32_t *vals, uint32_t n, uint32_t n_patterns,
                   uint64_t query, uint32_t *bits_out)
{
	RecipeTable t;
	build_recipe_table(k, keys, vals, n, n_patterns, t);
	if (bits_out) *bits_out = t.bits;
	const uint32_t mask = (1u << t.bits) - 1;
	for (uint32_t s = vg_h2b(vg_kmer_hash(query), t.bits);; s = (s + 1) & mask) {
		if (t.keys[s] == VG_EMPTY_KEY) return -1;
		if (t.keys[s] == query) return (int)t.vals[s];
	}
}

uint32_t sim_pack16(const uint8_t *c) { return pack16(c); }
uint32_t sim_rc32(uint32_t x, int L) { return vg_rc32(x, L); }

} // extern "C"
