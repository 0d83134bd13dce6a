// host_expand.cpp -- expansion of the compact control-matrix download (8 doubles per contact) into
// dense row-major 6x6 blocks (36 doubles).  Host data movement only; compiled by the host compiler
// (no CUDA).  See host_expand.h.
#include "host_expand.h"

#include <immintrin.h>

namespace blfccm {

void expand_ctrl_sse2(const double* src, double* dst, long long cnt)
{
    const __m128d z = _mm_setzero_pd();
    const bool nt = (reinterpret_cast<uintptr_t>(dst) & 15u) == 0;
    for (long long i = 0; i < cnt; ++i, src += 8, dst += 36) {
        const __m128d a = _mm_loadu_pd(src), b = _mm_loadu_pd(src + 2), c = _mm_loadu_pd(src + 4),
                      d = _mm_loadu_pd(src + 6);
        const __m128d v[18] = {
            _mm_unpacklo_pd(a, z), z, z,                          // gd 0 | 0 0 | 0 0
            _mm_unpacklo_pd(z, a), z, z,                          // 0 gd | 0 0 | 0 0
            z, _mm_unpacklo_pd(a, z), z,                          // 0 0 | gd 0 | 0 0
            z, _mm_unpackhi_pd(z, a), b,                          // 0 0 | 0 xx | xy xz
            z, _mm_unpacklo_pd(z, b), c,                          // 0 0 | 0 xy | yy yz
            z, _mm_unpackhi_pd(z, b), _mm_shuffle_pd(c, d, 1)};   // 0 0 | 0 xz | yz zz
        if (nt) {
            for (int j = 0; j < 18; ++j) _mm_stream_pd(dst + 2 * j, v[j]);
        } else {
            for (int j = 0; j < 18; ++j) _mm_storeu_pd(dst + 2 * j, v[j]);
        }
    }
    _mm_sfence();
}

// Two contacts = 72 doubles = nine full 64-byte lines: every line is one zeroing permute of a
// contact's compact vector {gd, xx, xy, xz, yy, yz, zz, pad} (the middle line takes its halves
// from both contacts), stored non-temporally as a whole line.
// dense index -> compact index:  0,7,14 <- 0 | 21 <- 1 | 22,27 <- 2 | 23,33 <- 3 | 28 <- 4 |
// 29,34 <- 5 | 35 <- 6; everything else zero.
__attribute__((target("avx512f"))) static void expand_ctrl_avx512(const double* src, double* dst,
                                                                   long long cnt)
{
    // line L of contact A covers dense entries 8L .. 8L+7 (L = 0..3), line 4 = A[32..35] | B[0..3],
    // lines 5..8 = B[4..35]
    const __m512i iA0 = _mm512_setr_epi64(0, 0, 0, 0, 0, 0, 0, 0);   // A[0..7]:   0->gd, 7->gd
    const __mmask8 mA0 = 0x81;
    const __m512i iA1 = _mm512_setr_epi64(0, 0, 0, 0, 0, 0, 0, 0);   // A[8..15]:  14->gd
    const __mmask8 mA1 = 0x40;
    const __m512i iA2 = _mm512_setr_epi64(0, 0, 0, 0, 0, 1, 2, 3);   // A[16..23]: 21,22,23
    const __mmask8 mA2 = 0xE0;
    const __m512i iA3 = _mm512_setr_epi64(0, 0, 0, 2, 4, 5, 0, 0);   // A[24..31]: 27,28,29
    const __mmask8 mA3 = 0x38;
    const __m512i iA4 = _mm512_setr_epi64(0, 3, 5, 6, 0, 0, 0, 0);   // A[32..35]: 33,34,35
    const __mmask8 mA4 = 0x0E;
    const __m512i iB4 = _mm512_setr_epi64(0, 0, 0, 0, 0, 0, 0, 0);   // B[0..3]:   0->gd  (lanes 4..7)
    const __mmask8 mB4 = 0x10;
    const __m512i iB5 = _mm512_setr_epi64(0, 0, 0, 0, 0, 0, 0, 0);   // B[4..11]:  7->gd
    const __mmask8 mB5 = 0x08;
    const __m512i iB6 = _mm512_setr_epi64(0, 0, 0, 0, 0, 0, 0, 0);   // B[12..19]: 14->gd
    const __mmask8 mB6 = 0x04;
    const __m512i iB7 = _mm512_setr_epi64(0, 1, 2, 3, 0, 0, 0, 2);   // B[20..27]: 21,22,23,27
    const __mmask8 mB7 = 0x8E;
    const __m512i iB8 = _mm512_setr_epi64(4, 5, 0, 0, 0, 3, 5, 6);   // B[28..35]: 28,29,33,34,35
    const __mmask8 mB8 = 0xE3;
    long long i = 0;
    for (; i + 2 <= cnt; i += 2, src += 16, dst += 72) {
        const __m512d A = _mm512_loadu_pd(src), B = _mm512_loadu_pd(src + 8);
        _mm512_stream_pd(dst + 0, _mm512_maskz_permutexvar_pd(mA0, iA0, A));
        _mm512_stream_pd(dst + 8, _mm512_maskz_permutexvar_pd(mA1, iA1, A));
        _mm512_stream_pd(dst + 16, _mm512_maskz_permutexvar_pd(mA2, iA2, A));
        _mm512_stream_pd(dst + 24, _mm512_maskz_permutexvar_pd(mA3, iA3, A));
        _mm512_stream_pd(dst + 32, _mm512_mask_permutexvar_pd(_mm512_maskz_permutexvar_pd(mA4, iA4, A),
                                                               mB4, iB4, B));
        _mm512_stream_pd(dst + 40, _mm512_maskz_permutexvar_pd(mB5, iB5, B));
        _mm512_stream_pd(dst + 48, _mm512_maskz_permutexvar_pd(mB6, iB6, B));
        _mm512_stream_pd(dst + 56, _mm512_maskz_permutexvar_pd(mB7, iB7, B));
        _mm512_stream_pd(dst + 64, _mm512_maskz_permutexvar_pd(mB8, iB8, B));
    }
    _mm_sfence();
    if (i < cnt) expand_ctrl_sse2(src, dst, cnt - i);   // odd tail
}

static bool have_avx512()
{
    static const bool v = __builtin_cpu_supports("avx512f");
    return v;
}

const char* expand_ctrl_isa() { return have_avx512() ? "avx512" : "sse2"; }

void expand_ctrl(const double* src, double* dst, long long cnt)
{
    // the AVX-512 form writes whole lines: destination 64-byte aligned (a contact pair is 576 bytes,
    // so the alignment carries over from pair to pair)
    if (cnt >= 2 && (reinterpret_cast<uintptr_t>(dst) & 63u) == 0 && have_avx512())
        expand_ctrl_avx512(src, dst, cnt);
    else
        expand_ctrl_sse2(src, dst, cnt);
}

}  // namespace blfccm
