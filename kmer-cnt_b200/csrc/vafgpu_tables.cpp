/*
 * vafgpu_tables.cpp -- see vafgpu_tables.hpp.
 *
 * Why anchors.  A pattern k-mer occupying stream offsets [p, e), e = p + k, covers the
 * aligned offset q = floor(e / S) * S and, because L <= k - S + 1, the whole L-mer
 * [q - L, q) that ENDS there.  So it is enough to look at one L-mer every S bases of the
 * stream: if it is the anchor some oriented pattern k-mer carries t = e - q bases before its
 * end (0 <= t < S), the k-mer ending at q + t is compared in full.  Every occurrence has
 * exactly one (q, t), so nothing is counted twice.  (Anchors end at aligned offsets, rather
 * than start there, so that a chunk of the stream only needs its LEFT neighbour, which the
 * kernel has already seen, never the next one, which may still be in flight.)
 * Both orientations of every canonical key are filed, which makes the forward k-mer of
 * the stream sufficient: canonical(x) is in the reference's map iff x or rc(x) is one of
 * its keys (vaf-counter.c:142-146,224,236).
 */
#include "vafgpu_tables.hpp"

#include <algorithm>
#include <cstdlib>
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
	items.reserve((size_t)n * 2 * S);
	seen.reserve((size_t)n * 2);
	for (uint32_t i = 0; i < n; ++i) {
		if (!seen.insert(keys[i]).second) continue; /* first insert wins */
		uint64_t f = ref_to_stream(keys[i], k), r = stream_revcomp(f, k);
		for (int orient = 0; orient < 2; ++orient) {
			uint64_t ok = orient ? r : f;
			if (orient && r == f) break; /* its own reverse complement (even k only) */
			for (int t = 0; t < S; ++t) { /* anchor = bases [k-t-L, k-t) of the oriented k-mer */
				uint32_t a = (uint32_t)(ok >> 2 * (k - t - L)) & amask;
				items.push_back({ok, vals[i], (uint32_t)t, a});
			}
		}
	}
	/* the distinct filter keys, without and with strand folding (sorted vectors: hash sets of 10^6
	 * keys cost most of a second here, which is start-up time of every run of the command line) */
	auto distinct = [&](int fold) {
		std::vector<uint32_t> v(items.size());
		for (size_t i = 0; i < items.size(); ++i) v[i] = fold ? vg_filter_key(items[i].anchor, L, 1) : items[i].anchor;
		std::vector<uint32_t> tmp(v.size()); /* least-significant-digit radix sort, four passes of eight bits */
		for (int shift = 0; shift < 32; shift += 8) {
			size_t count[257] = {0};
			for (uint32_t x : v) ++count[(x >> shift & 255u) + 1];
			for (int d = 0; d < 256; ++d) count[d + 1] += count[d];
			for (uint32_t x : v) tmp[count[x >> shift & 255u]++] = x;
			v.swap(tmp);
		}
		v.erase(std::unique(v.begin(), v.end()), v.end());
		return v;
	};
	const std::vector<uint32_t> plain = distinct(0);
	out.n_entries = (uint32_t)items.size();

	/* Filter: two bits per key in one 32-bit word.  A panel small enough to get >= 24 bits
	 * per key with both orientations filed keeps them (the kernel then hashes the forward
	 * anchor as it is); a large panel files strand-symmetric keys, halving the load of the
	 * filter at the price of a reverse complement per chunk in the kernel, and, its filter
	 * letting every tenth anchor through, is scanned by the deferred-lookup form of the kernel. */
	out.canon = (uint64_t)plain.size() * 24 > (uint64_t)VG_FILTER_BUDGET_WORDS(S, 0) * 32;
	out.defer = out.canon && VG_DEFER_OK(S);
	out.threads = VG_THREADS(S, out.defer);
	const uint32_t budget = VG_FILTER_BUDGET_WORDS(S, out.defer);
	const std::vector<uint32_t> folded = out.canon ? distinct(1) : std::vector<uint32_t>();
	const std::vector<uint32_t> &fkeys = out.canon ? folded : plain;
	out.n_filter_keys = (uint32_t)fkeys.size();
	uint64_t want = (uint64_t)out.n_filter_keys * 2;
	uint32_t nw = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(want, VG_MIN_FILTER_WORDS), budget);
	/* an odd word count: key -> hash * nw mod 2^32 (vg_hash_lo) is then a bijection, so every
	 * bit of it is usable further down the chain; the array is padded to whole 16-byte units */
	nw = (nw & ~3u) - 1u;
	out.filter_words = nw;
	out.filter.assign((nw + 3u) & ~3u, 0);
	for (uint32_t key : fkeys) out.filter[vg_filter_word(key, nw)] |= vg_filter_mask(key, nw);
	/* second level: 4 words (128 bits) per key, at least a page */
	const uint32_t nw2 = out.defer ? (uint32_t)std::max<uint64_t>(1024, (uint64_t)out.n_filter_keys * 4) : 4;
	out.filter2.assign(nw2, 0);
	if (out.defer)
		for (uint32_t key : fkeys) out.filter2[vg_filter2_word(key, nw, nw2)] |= vg_filter_mask(key, nw);

	/* exact table at <= 1/6 load, buckets of three tags: a full home bucket (a second L2
	 * round trip) is then rare.  Items are placed in the order they were generated, so a
	 * bucket's entries are consecutive in the payload array if we emit payloads bucket by
	 * bucket afterwards. */
	const uint32_t nb = (uint32_t)std::max<uint64_t>(64, (uint64_t)items.size() * 2);
	out.n_buckets = nb;
	std::vector<uint32_t> fill(nb, 0), where(items.size());
	std::vector<uint8_t> more(nb, 0);
	for (size_t i = 0; i < items.size(); ++i) {
		const uint32_t key = vg_filter_key(items[i].anchor, L, out.canon);
		uint32_t b = vg_bucket_home(vg_hash_lo(key, nw), nb);
		while (fill[b] == 3) {
			more[b] = 1;
			b = b + 1 == nb ? 0 : b + 1;
		}
		++fill[b];
		where[i] = b;
	}
	out.buckets.assign((size_t)(nb + 1) * 4, VG_FREE_TAG); /* + the empty bucket lanes without a survivor fetch */
	out.buckets[(size_t)nb * 4 + 3] = 0;
	out.slots.assign(items.size() ? items.size() : 1, vg_slot_t{VG_EMPTY_KEY, 0, 0});
	uint32_t base = 0;
	for (uint32_t b = 0; b < nb; ++b) {
		out.buckets[(size_t)b * 4 + 3] = base | (more[b] ? VG_CTRL_MORE : 0u);
		base += fill[b];
		fill[b] = 0; /* reused as the number of entries emitted so far */
	}
	for (size_t i = 0; i < items.size(); ++i) {
		const uint32_t b = where[i], pos = fill[b]++;
		out.buckets[(size_t)b * 4 + pos] = vg_tag(items[i].anchor, L);
		out.slots[(out.buckets[(size_t)b * 4 + 3] & ~VG_CTRL_MORE) + pos] = vg_slot_t{items[i].okey, items[i].val, items[i].off};
	}
}

} // namespace vafgpu
