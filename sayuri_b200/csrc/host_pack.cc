// Host-side packing of one position (InputData::planes, /root/reference/src/neural/network_basic.h:23-34) into the
// compact record the device unpacks (sb_packed_position, include/sayuri_b200.h).
//
// The encoder (src/neural/encoder.cc:80-100,296-319) produces 37 planes of {0,1} and 6 board-constant planes
// (rule, wave, +-komi/20, S/361, ones): every plane takes at most ONE non-zero value.  Such a plane is exactly
// (bit mask) x (that value), so the 62 KB fp32 record shrinks to 2.2 KB with no loss; the host->device traffic and
// the three host copies of the reference batcher (batch_forward_pipe.cc:9,178; cuda_forward_pipe.cc:694-701)
// go away.  Packing is EXACT or refused: a plane with two different non-zero values (or a NaN) makes
// sb_pack_position return 0 and the caller ships the raw fp32 planes for that sample instead.
#include <cstring>

#include "../../include/sayuri_b200.h"

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

// Scalar reference path: also the tail handler of the AVX2 path.
inline bool PackPlaneScalar(const float* p, int begin, int n, uint32_t* words, float& scale, bool& have) {
    for (int i = begin; i < n; ++i) {
        const float v = p[i];
        if (v != 0.0f) {               // NaN != 0 is true and fails the equality below: refused
            if (!have) {
                scale = v;
                have = true;
            }
            if (!(v == scale)) return false;
            words[i >> 5] |= 1u << (i & 31);
        }
    }
    return true;
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) bool PackPlaneAvx2(const float* p, int n, uint32_t* words, float& scale) {
    bool have = false;
    const __m256 zero = _mm256_setzero_ps();
    __m256 vs = zero;
    int i = 0;
    for (; i + 32 <= n; i += 32) {
        uint32_t word = 0;
        uint32_t bad = 0;
#pragma GCC unroll 4
        for (int k = 0; k < 4; ++k) {
            const __m256 v = _mm256_loadu_ps(p + i + 8 * k);
            const uint32_t nz = (uint32_t)_mm256_movemask_ps(_mm256_cmp_ps(v, zero, _CMP_NEQ_UQ));
            if (nz && !have) {
                scale = p[i + 8 * k + __builtin_ctz(nz)];
                vs = _mm256_set1_ps(scale);
                have = true;
            }
            const uint32_t eq = (uint32_t)_mm256_movemask_ps(_mm256_cmp_ps(v, vs, _CMP_EQ_OQ));
            bad |= nz & ~eq;
            word |= nz << (8 * k);
        }
        if (bad) return false;
        words[i >> 5] = word;
    }
    return PackPlaneScalar(p, i, n, words, scale, have);
}
#endif

bool HaveAvx2() {
#if defined(__x86_64__)
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
#else
    return false;
#endif
}

}  // namespace

extern "C" int sb_pack_position(const float* planes, int board_size, int offset, sb_packed_position* out) {
    if (!planes || !out || board_size < 2 || board_size > SB_MAX_BOARD_SIZE) return 0;
    const int n = board_size * board_size;
    std::memset(out->bits, 0, sizeof(out->bits));
    out->board_size = board_size;
    out->offset = offset;
    out->flags = 0;
    out->reserved = 0;
    const bool avx2 = HaveAvx2();
    for (int c = 0; c < SB_INPUT_CHANNELS; ++c) {
        float scale = 0.0f;
        bool ok;
#if defined(__x86_64__)
        if (avx2) {
            ok = PackPlaneAvx2(planes + (size_t)c * n, n, out->bits[c], scale);
        } else
#endif
        {
            bool have = false;
            ok = PackPlaneScalar(planes + (size_t)c * n, 0, n, out->bits[c], scale, have);
        }
        (void)avx2;
        if (!ok) {
            out->flags = SB_PACKED_RAW;
            return 0;
        }
        out->scale[c] = scale;
    }
    return 1;
}

// Inverse (host), for tests: expand a packed record back into fp32 planes at the native board size.
extern "C" int sb_unpack_position(const sb_packed_position* rec, float* planes) {
    if (!rec || !planes || (rec->flags & SB_PACKED_RAW)) return 0;
    const int n = rec->board_size * rec->board_size;
    for (int c = 0; c < SB_INPUT_CHANNELS; ++c)
        for (int i = 0; i < n; ++i) planes[(size_t)c * n + i] = ((rec->bits[c][i >> 5] >> (i & 31)) & 1u) ? rec->scale[c] : 0.0f;
    return 1;
}
