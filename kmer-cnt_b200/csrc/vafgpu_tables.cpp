/*
 * vafgpu_tables.cpp -- see vafgpu_tables.hpp.
 *
 * Why anchors.  A pattern k-mer occurring at stream offset p covers the aligned offset
 * q = ceil(p / S) * S and, because L <= k - S + 1, the whole anchor [q, q + L).  So it is
 * enough to look at one L-mer every S bases of the stream: if it is the anchor some
 * oriented pattern k-mer carries at offset o = q - p (0 <= o < S), the k-mer at q - o is
 * compared in full.  Every occurrence has exactly one (q, o), so nothing is counted twice.
 * Both orientations of every canonical key are filed, which makes the forward k-mer of
 * the stream sufficient: canonical(x) is in the reference's map iff x or rc(x) is one of
 * its keys (vaf-counter.c:142-146,224,236).
 */
#include "vafgpu_tables.hpp"

#include <algorithm>
#include <unordered_set>

namespace vafgpu {

Plan make_plan(int k)
{
	Plan p;
	p.k = k;
	if (k >= 27) p.stride = 16;
	else if (k >= 19) p.stride = 8;
	else if (k >= 15) p.stride = 4;
	else if (k >= 13) p.stride = 2;
	else p.stride = 1;
	p.len = std::min(16, k - p.stride + 1);
	return p;
}

static uint32_t khashl_bits(uint32_t want) /* khashl.h:152-160 */
{
	uint32_t j = 0, x = want;
	while ((x >>= 1) != 0) ++j;
	if (want & (want - 1)) ++j;
	return j > 2 ? j : 2;
}

void build_recipe_table(int k, const uint64_t *keys, const uint32_t *vals, uint32_t n,
                        uint32_t n_patterns, RecipeTable &out)
{
	(void)k;
	uint32_t bits = khashl_bits(n_patterns * 3u);
	while ((uint64_t)n * 4 > (3ull << bits)) ++bits; /* never above 75 % load, khashl.h:202 */
	out.bits = bits;
	out.keys.assign((size_t)1 << bits, VG_EMPTY_KEY);
	out.vals.assign((size_t)1 << bits, 0);
	const uint32_t mask = (1u << bits) - 1;
	for (uint32_t i = 0; i < n; ++i) {
		uint32_t b = vg_h2b(vg_kmer_hash(keys[i]), bits);
		while (out.keys[b] != VG_EMPTY_KEY && out.keys[b] != keys[i]) b = (b + 1) & mask;
		if (out.keys[b] == VG_EMPTY_KEY) { /* first insert wins, vaf-counter.c:226-230 */
			out.keys[b] = keys[i];
			out.vals[b] = vals[i];
		}
	}
}

uint64_t ref_to_stream(uint64_t ref_key, int k)
{
	uint64_t o = 0;
	for (int i = 0; i < k; ++i) {
		uint64_t c = (ref_key >> 2 * (k - 1 - i)) & 3; /* base i, reference code */
		o |= (c ^ (c >> 1)) << 2 * i;                 /* A0 C1 G2 T3 -> A0 C1 G3 T2 */
	}
	return o;
}

uint64_t stream_revcomp(uint64_t okey, int k)
{
	uint64_t r = 0;
	for (int i = 0; i < k; ++i) r |= (((okey >> 2 * i) & 3) ^ 2) << 2 * (k - 1 - i);
	return r;
}

void build_anchor_tables(int k, const uint64_t *keys, const uint32_t *vals, uint32_t n,
                         AnchorTables &out)
{
	const Plan plan = make_plan(k);
	const int S = plan.stride, L = plan.len;
	const uint32_t amask = vg_mask32(L);
	out.plan = plan;

	struct Item { uint64_t okey; uint32_t val, off, anchor; };
	std::vector<Item> items;
	std::unordered_set<uint64_t> seen;
	std::unordered_set<uint32_t> canon;
	items.reserve((size_t)n * 2 * S);
	for (uint32_t i = 0; i < n; ++i) {
		if (!seen.insert(keys[i]).second) continue; /* first insert wins */
		uint64_t f = ref_to_stream(keys[i], k), r = stream_revcomp(f, k);
		for (int orient = 0; orient < 2; ++orient) {
			uint64_t ok = orient ? r : f;
			if (orient && r == f) break; /* its own reverse complement (even k only) */
			for (int o = 0; o < S; ++o) {
				uint32_t a = (uint32_t)(ok >> 2 * o) & amask;
				items.push_back({ok, vals[i], (uint32_t)o, a});
				canon.insert(vg_canon32(a, L));
			}
		}
	}
	out.n_entries = (uint32_t)items.size();
	out.n_filter_keys = (uint32_t)canon.size();

	/* filter: 64 bits per distinct anchor if shared memory allows, two bits set per anchor */
	uint64_t want = (uint64_t)out.n_filter_keys * 2;
	uint32_t nw = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(want, VG_MIN_FILTER_WORDS),
	                                           VG_MAX_FILTER_WORDS);
	nw = (nw + 3u) & ~3u;
	out.filter.assign(nw, 0);
	for (uint32_t c : canon) {
		uint32_t h = vg_filter_hash(c);
		out.filter[vg_filter_word(h, nw)] |= vg_filter_mask(h);
	}

	/* exact table at <= 50 % load */
	uint32_t bits = 4;
	while ((1ull << bits) < (uint64_t)items.size() * 2) ++bits;
	out.slot_bits = bits;
	out.slots.assign((size_t)1 << bits, vg_slot_t{VG_EMPTY_KEY, 0, 0});
	const uint32_t smask = (1u << bits) - 1;
	for (const Item &it : items) {
		uint32_t s = vg_slot_home(it.anchor, bits);
		while (out.slots[s].okey != VG_EMPTY_KEY) s = (s + 1) & smask;
		out.slots[s] = vg_slot_t{it.okey, it.val, it.off};
	}
}

} // namespace vafgpu
