/*
 * vafgpu_pack.cpp -- host side of the byte rules: canonicalise a read to {A,C,G,T,N} while
 * it is copied into a staging block, so that the device sees exactly the bases the
 * reference's encoder sees (vaf-counter.c:261-291: PSHUFB low-nibble rule for offsets below
 * len & ~15, strict table vaf-counter.c:73-90 for the tail).  Runs at memcpy speed: one
 * PSHUFB per 16 bytes where SSSE3 is available, a table walk otherwise.
 */
#include <cstddef>
#include <cstring>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/vafgpu.h"

namespace {

const char kNibbleRule[16 + 1] = "NANCTTNGNNNNNNNN"; /* vaf-counter.c:272-275 */

struct StrictRule {
	char t[256];
	StrictRule()
	{
		memset(t, 'N', sizeof t);
		t[0] = 'A', t[1] = 'C', t[2] = 'G', t[3] = 'T'; /* vaf-counter.c:74 */
		t['A'] = t['a'] = 'A';
		t['C'] = t['c'] = 'C';
		t['G'] = t['g'] = 'G';
		t['T'] = t['t'] = t['U'] = t['u'] = 'T';
	}
};
const StrictRule kStrictRule;

#if defined(__x86_64__)
__attribute__((target("ssse3"))) void nibble_rule_ssse3(const char *seq, size_t n16, char *out)
{
	const __m128i lut = _mm_loadu_si128(reinterpret_cast<const __m128i *>(kNibbleRule));
	const __m128i low = _mm_set1_epi8(0x0F);
	for (size_t i = 0; i < n16; i += 16) {
		__m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i *>(seq + i));
		_mm_storeu_si128(reinterpret_cast<__m128i *>(out + i), _mm_shuffle_epi8(lut, _mm_and_si128(v, low)));
	}
}
#endif

} // namespace

extern "C" void vafgpu_canonicalise_read(const char *seq, size_t len, char *out, int simd_rule)
{
	const size_t body = simd_rule ? (len & ~(size_t)15) : 0; /* vaf-counter.c:278 */
	size_t i = 0;
#if defined(__x86_64__)
	static const bool have_ssse3 = __builtin_cpu_supports("ssse3");
	if (have_ssse3 && body) {
		nibble_rule_ssse3(seq, body, out);
		i = body;
	}
#endif
	for (; i < body; ++i) out[i] = kNibbleRule[(unsigned char)seq[i] & 15];
	for (; i < len; ++i) out[i] = kStrictRule.t[(unsigned char)seq[i]]; /* vaf-counter.c:288-290 */
}
