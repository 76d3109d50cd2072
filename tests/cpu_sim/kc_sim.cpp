/*
 * kc_sim.cpp -- TEST-ONLY host emulation of the counting mode's bookkeeping: the table / list /
 * inbox geometry of csrc/kcgpu_kernels.cuh and the owner-region-tag split the kernels apply,
 * driven by the same header (compiled for the host), so that the arithmetic is checked
 * without a GPU: every k, 1..16 owners, tables from 2^12 to 2^36 slots.
 *
 * sim_kc_count runs a stream through push -> route -> flush with host loops that follow the
 * kernels' bookkeeping step by step (inbox, region lists, table) and returns the 256-bin
 * histogram over all owners.
 *
 * sim_kc_tile_run checks where the entries of one region's run of a sorted tile are sent
 * (kc_tile_word / kc_tile_fits of the header) against the plain statement of it.
 *
 * sim_kc_extract runs the tile kernels' own extraction (kc_extract16 of the header: packed
 * 48-byte windows, no byte loop) over a stream, chunk by chunk as the threads do.
 */
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../kmer-cnt_b200/csrc/kcgpu_kernels.cuh"

using namespace kcgpu;

namespace {

struct Owner {
	std::vector<uint64_t> mem; /* table | lists | cursors, as on the device */
	uint64_t n_slots, list_cap;
	uint32_t region_bits, rslot_bits;
	uint64_t *base() { return mem.data(); }
};

bool is_base(uint8_t b)
{
	b &= 0xDF;
	return b == 'A' || b == 'C' || b == 'G' || b == 'T' || b == 'U';
}

/* kc_insert_from on the host */
bool insert(Owner &o, uint64_t q)
{
	const uint64_t region = q & ((1ull << o.region_bits) - 1), tag = q >> o.region_bits;
	if (tag >> KC_TAG_BITS) return false; /* would not fit beside the count */
	uint64_t *slice = o.base() + (region << o.rslot_bits);
	const uint64_t rmask = (1ull << o.rslot_bits) - 1;
	uint64_t pos = (tag * 0x9E3779B97F4A7C15ull) >> (64 - o.rslot_bits);
	for (uint64_t tries = 0; tries <= rmask; ++tries, pos = (pos + 1) & rmask) {
		uint64_t &v = slice[pos];
		if (v == 0) {
			v = (tag + 1) << KC_COUNT_BITS | 1;
			return true;
		}
		if (v >> KC_COUNT_BITS == tag + 1) {
			if ((v & KC_COUNT_MAX) < KC_COUNT_MAX) ++v;
			return true;
		}
	}
	return false;
}

} // namespace

extern "C" {

/* geometry of one allocation: out = {region_bits(k), alloc bytes, lists offset, cursors offset,
 * inbox cursor offset, inbox capacity, end of the region lists (several owners)} in bytes / entries */
void sim_kc_geometry(int k, uint64_t n_slots, uint64_t list_cap, uint32_t region_bits, uint64_t *out)
{
	uint64_t *base = nullptr;
	out[0] = kc_region_bits(k);
	out[1] = kc_alloc_bytes(n_slots, list_cap, region_bits);
	out[2] = (uint64_t)((char *)kc_lists_of(base, n_slots) - (char *)base);
	out[3] = (uint64_t)((char *)kc_cursors_of(base, n_slots, list_cap, region_bits) - (char *)base);
	out[4] = (uint64_t)((char *)kc_inbox_cursor(base, n_slots, list_cap, region_bits) - (char *)base);
	out[5] = kc_inbox_cap(list_cap, region_bits);
	out[6] = out[2] + (out[5] + (list_cap / 2 << region_bits)) * 8;
}

/* one region's run of a tile: entries lbase .. lbase + c - 1, cursor value g; returns 0 when every entry is sent where
 * it belongs (kc_tile_word / kc_tile_fits), else 1 + the first entry that is not */
uint64_t sim_kc_tile_run(uint64_t g, uint32_t c, uint32_t lbase, uint32_t region, uint64_t cap, uint64_t stride)
{
	const unsigned long long word = kc_tile_word(g, c, lbase, region, cap, stride);
	for (uint32_t i = lbase; i < lbase + c; ++i) {
		uint64_t where;
		const bool fits = kc_tile_fits(word, i, &where);
		const uint64_t pos = g + (i - lbase);
		if (g + c <= cap) {
			if (!fits || where != (uint64_t)region * stride + pos) return 1 + i;
		} else {
			if (fits || where != pos) return 1 + i;
		}
	}
	return 0;
}

uint64_t sim_kc_hash64(uint64_t key, int k) { return kc_hash64(key, (1ull << 2 * k) - 1); }

/* hash64 of every canonical k-mer of the stream (n_bytes a multiple of 16), in stream order, as the
 * threads of kc_scan_tile_kernel / kc_push_tile_kernel compute them; returns how many (out may be
 * NULL to count only, else it takes up to cap of them) */
uint64_t sim_kc_extract(int k, const uint8_t *bytes, uint64_t n_bytes, uint64_t *out, uint64_t cap)
{
	std::vector<uint4> chunks(n_bytes / 16);
	memcpy(chunks.data(), bytes, chunks.size() * 16);
	const Extract x = kc_extract_of(k);
	uint64_t n = 0;
	for (uint64_t c = 0; c < chunks.size(); ++c) {
		uint64_t h[KC_TILE_N];
		const uint32_t ok = kc_extract16(chunks.data(), c, chunks.size(), x, h);
		for (int j = 0; j < KC_TILE_N; ++j)
			if (ok >> j & 1) {
				if (out && n < cap) out[n] = h[j];
				++n;
			}
	}
	return n;
}

/* push -> route -> flush over n_parts owners with tables of 2^table_bits slots and list_cap
 * entries per region; returns the number of k-mers that could not be placed (0 expected) */
uint64_t sim_kc_count(int k, int n_parts, uint32_t table_bits, uint64_t list_cap, const uint8_t *bytes, uint64_t n_bytes,
                      uint64_t hist[256], uint64_t *n_direct_out)
{
	const uint32_t region_bits = kc_region_bits(k) > 3 ? kc_region_bits(k) : 3;
	std::vector<Owner> own((size_t)n_parts);
	for (Owner &o : own) {
		o.n_slots = 1ull << table_bits;
		o.list_cap = list_cap;
		o.region_bits = region_bits;
		o.rslot_bits = table_bits - region_bits;
		o.mem.assign(kc_alloc_bytes(o.n_slots, list_cap, region_bits) / 8, 0);
	}
	uint64_t lost = 0, n_direct = 0;
	const uint64_t mask = (1ull << 2 * k) - 1;
	uint64_t fw = 0, rv = 0;
	int run = 0;
	auto flush = [&](Owner &o) { /* kc_route_kernel then kc_flush_kernel */
		uint64_t *lists = kc_lists_of(o.base(), o.n_slots);
		unsigned long long *cur = kc_cursors_of(o.base(), o.n_slots, o.list_cap, o.region_bits);
		unsigned long long &inbox_n = *kc_inbox_cursor(o.base(), o.n_slots, o.list_cap, o.region_bits);
		const uint64_t icap = kc_inbox_cap(o.list_cap, o.region_bits), cap2 = o.list_cap / 2;
		uint64_t *lists2 = lists + icap;
		const uint64_t n = inbox_n < icap ? inbox_n : icap;
		for (uint64_t i = 0; i < n; ++i) {
			const uint64_t q = lists[i], region = q & ((1ull << o.region_bits) - 1);
			const uint64_t at = cur[region * KC_CURSOR_STRIDE]++;
			if (at < cap2) lists2[region * cap2 + at] = q;
			else ++n_direct, lost += !insert(o, q);
		}
		for (uint64_t r = 0; r < (1ull << o.region_bits); ++r) {
			const uint64_t filled = cur[r * KC_CURSOR_STRIDE], m = filled < cap2 ? filled : cap2;
			for (uint64_t i = 0; i < m; ++i) lost += !insert(o, lists2[r * cap2 + i]);
			cur[r * KC_CURSOR_STRIDE] = 0;
		}
		inbox_n = 0;
	};
	for (uint64_t i = 0; i < n_bytes; ++i) {
		const uint8_t b = bytes[i];
		if (!is_base(b)) {
			run = 0;
			continue;
		}
		uint64_t c = (b >> 1) & 3;
		c ^= c >> 1;
		fw = (fw << 2 | c) & mask;
		rv = rv >> 2 | (3 - c) << 2 * (k - 1);
		if (++run < k) continue;
		const uint64_t h = kc_hash64(fw < rv ? fw : rv, mask);
		const uint64_t q = h / (uint64_t)n_parts;
		Owner &o = own[h % (uint64_t)n_parts];
		if (q * (uint64_t)n_parts + h % (uint64_t)n_parts != h) ++lost;
		unsigned long long &inbox_n = *kc_inbox_cursor(o.base(), o.n_slots, o.list_cap, o.region_bits);
		const uint64_t at = inbox_n++;
		if (at < kc_inbox_cap(o.list_cap, o.region_bits)) kc_lists_of(o.base(), o.n_slots)[at] = q;
		else ++n_direct, lost += !insert(o, q);
		if (i % 40000 == 39999) /* a flush now and then, as the host does when the lists are due */
			for (Owner &x : own) flush(x);
	}
	for (Owner &x : own) flush(x);
	memset(hist, 0, 256 * sizeof hist[0]);
	for (Owner &o : own)
		for (uint64_t s = 0; s < o.n_slots; ++s)
			if (o.mem[s]) {
				const uint32_t cnt = (uint32_t)o.mem[s] & KC_COUNT_MAX;
				hist[cnt < 255 ? cnt : 255]++;
			}
	if (n_direct_out) *n_direct_out = n_direct;
	return lost;
}
}
