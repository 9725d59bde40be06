// k1_moments.cuh -- K1: streaming intensity statistics, one WARP per (object, channel) tile.
//
// Replaces, per channel (NB = reference notebook raw line):
//   min/max  NB:241,251   total NB:254   mean NB:257   std NB:258   kurtosis NB:259   skew NB:260
//
// Arithmetic: pass 1 gets min / max / sum with packed 16-bit SIMD (2 pixels per instruction);
// pass 2 accumulates sum y^2, y^3, y^4 of y = x - p (p = integer nearest the mean) as
// integer-valued doubles (y, y^2 and sum y^2 are exact), then shifts the pivot to the exact
// mean analytically.  For n <= 4096 the tile lives in registers between the passes, so HBM is
// read exactly once (16 x LDG.128 in flight per lane).
#pragma once
#include "common.cuh"

namespace imfeat {

constexpr int kK1Vec = 16;                 // 16 vectors x 8 px x 32 lanes = 4096 px in registers
constexpr double kBias52 = 4503599627370496.0;            // 2^52
constexpr double kBiasC0 = 4503599627370496.0 + 1048576.0;  // 2^52 + 2^20

__device__ __forceinline__ void k1_p1_word(uint32_t w, uint32_t& mn2, uint32_t& mx2, uint32_t& sum) {
    mn2 = __vminu2(mn2, w);
    mx2 = __vmaxu2(mx2, w);
    sum = __dp2a_lo(w, 0x0101u, sum);
}

// L = x + (2^20 - p) (always in [0, 2^21)); double(2^52 + L) - (2^52 + 2^20) == x - p exactly.
__device__ __forceinline__ void k1_p2_px(uint32_t L, double& S2, double& S3, double& S4) {
    const double y = __hiloint2double(0x43300000, (int)L) - kBiasC0;
    const double y2 = y * y;
    S2 += y2;
    S3 = fma(y2, y, S3);
    S4 = fma(y2, y2, S4);
}

__device__ __forceinline__ void k1_p2_word(uint32_t w, uint32_t K, double* S) {
    k1_p2_px((w & 0xffffu) + K, S[0], S[1], S[2]);
    k1_p2_px((w >> 16) + K, S[3], S[4], S[5]);
}

// Expand 4 mask bytes (non-zero = inside) to two words of 16-bit lane masks.
__device__ __forceinline__ void mask_halfwords(uint32_t m4, uint32_t& h01, uint32_t& h23) {
    const uint32_t c = __vcmpne4(m4, 0u);  // 0xff per non-zero byte
    h01 = __byte_perm(c, 0u, 0x1100);
    h23 = __byte_perm(c, 0u, 0x3322);
}

struct K1Acc {
    uint32_t mn2, mx2, sum, cnt;
};

__device__ __forceinline__ void k1_epilogue(const Params& P, const Tile& T, uint32_t n_eff,
                                            uint32_t vmin, uint32_t vmax, uint32_t total,
                                            long long p, double S2, double S3, double S4) {
    double* o = T.out_row + P.col_basic + kNBasic * T.slot;
    if (n_eff == 0) {
        const double nan = qnan();
        o[0] = nan; o[10] = nan; o[11] = nan; o[12] = nan; o[13] = nan; o[14] = nan; o[15] = nan;
        if (T.status) atomicOr(T.status, kStEmptyMask);
        return;
    }
    const double nn = (double)n_eff;
    const double mean = (double)total / nn;
    const double r = (double)((long long)total - p * (long long)n_eff);  // sum of y, |r| <= n/2
    const double d = r / nn;                                             // mean - p
    const double C2 = S2 - r * r / nn;
    const double C3 = S3 - 3.0 * d * S2 + 2.0 * r * d * d;
    const double C4 = S4 - 4.0 * d * S3 + 6.0 * d * d * S2 - 3.0 * r * d * d * d;
    const double m2 = C2 / nn, m3 = C3 / nn, m4 = C4 / nn;
    o[0] = (double)vmin;
    o[10] = (double)vmax;
    o[11] = (double)total;
    o[12] = mean;
    o[13] = sqrt(m2);
    const double thr = 2.220446049250313e-16 * mean;   // scipy: m2 <= (eps*mean)^2 -> NaN
    if (m2 <= thr * thr) {
        o[14] = qnan();
        o[15] = qnan();
        if (T.status) atomicOr(T.status, kStConstant);
    } else {
        o[14] = m4 / (m2 * m2) - 3.0;
        o[15] = m3 / (m2 * sqrt(m2));
    }
}

template <bool MASKED>
__global__ void __launch_bounds__(256) k1_moments_kernel(const __grid_constant__ Params P) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;

    for (long long t = warp0; t < P.n_tiles; t += nwarps) {
        const Tile T = resolve_tile(P, t);
        const uint4* px4 = reinterpret_cast<const uint4*>(T.px);
        const int nfull = T.n >> 3, rem = T.n & 7;
        uint32_t mn2 = 0xffffffffu, mx2 = 0u, sum = 0u, cnt = 0u;
        double S[6] = {0, 0, 0, 0, 0, 0};
        long long p = 0;

        if (!MASKED && T.n <= kK1Vec * 256) {
            // ---- register-resident fast path: one HBM read ----
            uint4 v[kK1Vec];
#pragma unroll
            for (int i = 0; i < kK1Vec; ++i) {
                const int idx = lane + 32 * i;
                if (idx < nfull) v[i] = ld_stream(px4 + idx);
            }
            uint32_t xt = 0;
            if (lane < rem) xt = T.px[nfull * 8 + lane];
#pragma unroll
            for (int i = 0; i < kK1Vec; ++i) {
                if (lane + 32 * i < nfull) {
                    k1_p1_word(v[i].x, mn2, mx2, sum);
                    k1_p1_word(v[i].y, mn2, mx2, sum);
                    k1_p1_word(v[i].z, mn2, mx2, sum);
                    k1_p1_word(v[i].w, mn2, mx2, sum);
                }
            }
            if (lane < rem) { mn2 = __vminu2(mn2, xt | 0xffff0000u); mx2 = __vmaxu2(mx2, xt); sum += xt; }
            const uint32_t total = __reduce_add_sync(0xffffffffu, sum);
            p = ((long long)total + (T.n >> 1)) / T.n;
            const uint32_t K = (1u << 20) - (uint32_t)p;
#pragma unroll
            for (int i = 0; i < kK1Vec; ++i) {
                if (lane + 32 * i < nfull) {
                    k1_p2_word(v[i].x, K, S);
                    k1_p2_word(v[i].y, K, S);
                    k1_p2_word(v[i].z, K, S);
                    k1_p2_word(v[i].w, K, S);
                }
            }
            if (lane < rem) k1_p2_px(xt + K, S[0], S[1], S[2]);
            cnt = 0;  // n_eff = T.n below
            sum = total;
        } else {
            // ---- generic two-pass path (large tiles and masked tiles); pass 2 re-reads L1/L2 ----
            const uint2* mk2 = reinterpret_cast<const uint2*>(T.mk);
            for (int idx = lane; idx < nfull; idx += 32) {
                uint4 v = ld_reuse(px4 + idx);
                if (MASKED) {
                    const uint2 m = __ldg(mk2 + idx);
                    uint32_t h0, h1, h2, h3;
                    mask_halfwords(m.x, h0, h1);
                    mask_halfwords(m.y, h2, h3);
                    cnt += (__popc(h0) + __popc(h1) + __popc(h2) + __popc(h3)) >> 4;
                    mn2 = __vminu2(mn2, v.x | ~h0); mn2 = __vminu2(mn2, v.y | ~h1);
                    mn2 = __vminu2(mn2, v.z | ~h2); mn2 = __vminu2(mn2, v.w | ~h3);
                    v.x &= h0; v.y &= h1; v.z &= h2; v.w &= h3;
                    mx2 = __vmaxu2(mx2, v.x); mx2 = __vmaxu2(mx2, v.y);
                    mx2 = __vmaxu2(mx2, v.z); mx2 = __vmaxu2(mx2, v.w);
                    sum = __dp2a_lo(v.x, 0x0101u, sum); sum = __dp2a_lo(v.y, 0x0101u, sum);
                    sum = __dp2a_lo(v.z, 0x0101u, sum); sum = __dp2a_lo(v.w, 0x0101u, sum);
                } else {
                    k1_p1_word(v.x, mn2, mx2, sum);
                    k1_p1_word(v.y, mn2, mx2, sum);
                    k1_p1_word(v.z, mn2, mx2, sum);
                    k1_p1_word(v.w, mn2, mx2, sum);
                }
            }
            if (lane < rem) {
                const uint32_t xt = T.px[nfull * 8 + lane];
                const bool ok = !MASKED || T.mk[nfull * 8 + lane] != 0;
                if (ok) { mn2 = __vminu2(mn2, xt | 0xffff0000u); mx2 = __vmaxu2(mx2, xt); sum += xt; cnt += 1; }
            }
            const uint32_t total = __reduce_add_sync(0xffffffffu, sum);
            const uint32_t n_eff = MASKED ? __reduce_add_sync(0xffffffffu, cnt) : (uint32_t)T.n;
            p = n_eff ? ((long long)total + (n_eff >> 1)) / n_eff : 0;
            const uint32_t K = (1u << 20) - (uint32_t)p;
            for (int idx = lane; idx < nfull; idx += 32) {
                const uint4 v = ld_reuse(px4 + idx);
                if (MASKED) {
                    const uint2 m = __ldg(mk2 + idx);
                    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t mb = (k < 2 ? m.x : m.y) >> (16 * (k & 1));
                        if (mb & 0xffu) k1_p2_px((w[k] & 0xffffu) + K, S[0], S[1], S[2]);
                        if (mb & 0xff00u) k1_p2_px((w[k] >> 16) + K, S[3], S[4], S[5]);
                    }
                } else {
                    k1_p2_word(v.x, K, S);
                    k1_p2_word(v.y, K, S);
                    k1_p2_word(v.z, K, S);
                    k1_p2_word(v.w, K, S);
                }
            }
            if (lane < rem) {
                const uint32_t xt = T.px[nfull * 8 + lane];
                const bool ok = !MASKED || T.mk[nfull * 8 + lane] != 0;
                if (ok) k1_p2_px(xt + K, S[0], S[1], S[2]);
            }
            cnt = n_eff;
            sum = total;
        }

        // ---- warp reduction + epilogue ----
        const uint32_t vmin = __reduce_min_sync(0xffffffffu, min(mn2 & 0xffffu, mn2 >> 16));
        const uint32_t vmax = __reduce_max_sync(0xffffffffu, max(mx2 & 0xffffu, mx2 >> 16));
        const double S2 = warp_sum(S[0] + S[3]);
        const double S3 = warp_sum(S[1] + S[4]);
        const double S4 = warp_sum(S[2] + S[5]);
        const uint32_t n_eff = (!MASKED && T.n <= kK1Vec * 256) ? (uint32_t)T.n : cnt;
        if (lane == 0) k1_epilogue(P, T, n_eff, vmin, vmax, sum, p, S2, S3, S4);
    }
}

}  // namespace imfeat
