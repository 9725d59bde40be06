// k1_moments.cuh -- K1: streaming intensity statistics, one WARP per (object, channel) tile.
//
// Replaces, per channel (NB = reference notebook raw line):
//   min/max  NB:241,251   total NB:254   mean NB:257   std NB:258   kurtosis NB:259   skew NB:260
//
// Single pass over HBM.  min / max / sum use packed 16-bit SIMD (2 pixels per instruction).  The
// central moments are accumulated around an integer pivot p taken from a small sample of the
// tile (the mean of ~256 valid pixels spread over it), y = x - p: as exact 64-bit integers (IMAD.WIDE)
// while |y| <= 11,585 -- all 12-bit data -- and otherwise (sample range too wide, or the limit found
// broken after the pass) as doubles
// (y formed exactly by an integer add into the mantissa of 2^52, then sum y^2 (exact), y^3, y^4
// with FP64 FMAs).
// The pivot is shifted to the exact mean analytically in the epilogue; because the pivot is the
// mean of a subset of the pixels, |mean - p| is bounded by a small multiple of the standard
// deviation and the shift is numerically benign.
#pragma once
#include "common.cuh"

namespace imfeat {

constexpr double kBiasC0 = 4503599627370496.0 + 1048576.0;  // 2^52 + 2^20
constexpr int kK1Unroll = 4;

// double(2^52 + 2^20 + (x - p)) built in one IMAD.WIDE: bits = x + ((0x43300000 << 32) | (2^20 - p))
__device__ __forceinline__ double k1_biased(uint32_t x, unsigned long long c64) {
    unsigned long long bits;
    asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(bits) : "r"(x), "l"(c64));
    return __longlong_as_double((long long)bits);
}
__device__ __forceinline__ void k1_px(uint32_t x, unsigned long long c64, double& S2, double& S3, double& S4) {
    const double y = k1_biased(x, c64) - kBiasC0;      // x - p, exact
    const double y2 = y * y;                           // exact (< 2^34)
    S2 += y2;                                          // exact (< 2^53)
    S3 = fma(y2, y, S3);
    S4 = fma(y2, y2, S4);
}

// Expand 4 mask bytes (non-zero = inside) to two words of 16-bit lane masks.
__device__ __forceinline__ void mask_halfwords(uint32_t m4, uint32_t& h01, uint32_t& h23) {
    const uint32_t c = __vcmpne4(m4, 0u);  // 0xff per non-zero byte
    h01 = __byte_perm(c, 0u, 0x1100);
    h23 = __byte_perm(c, 0u, 0x3322);
}

struct K1State {
    uint32_t mn2, mx2, sum, cnt;
    double S[6];
};

template <bool MASKED>
__device__ __forceinline__ void k1_vec(const uint4& v, const uint2& m, unsigned long long c64, K1State& st) {
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
    if (MASKED) {
        uint32_t h[4];
        mask_halfwords(m.x, h[0], h[1]);
        mask_halfwords(m.y, h[2], h[3]);
        st.cnt += (__popc(h[0]) + __popc(h[1]) + __popc(h[2]) + __popc(h[3])) >> 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            st.mn2 = __vminu2(st.mn2, w[k] | ~h[k]);
            w[k] &= h[k];
            st.mx2 = __vmaxu2(st.mx2, w[k]);
            st.sum = __dp2a_lo(w[k], 0x0101u, st.sum);
            // branch-free: a pixel outside the mask (already zeroed) gets the pivot-free bias, so
            // its y is exactly 0 and adds nothing to the three sums
            const unsigned long long cz = (c64 & 0xffffffff00000000ull) | (1ull << 20);
            k1_px(w[k] & 0xffffu, (h[k] & 0xffffu) ? c64 : cz, st.S[0], st.S[1], st.S[2]);
            k1_px(w[k] >> 16, (h[k] >> 16) ? c64 : cz, st.S[3], st.S[4], st.S[5]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            st.mn2 = __vminu2(st.mn2, w[k]);
            st.mx2 = __vmaxu2(st.mx2, w[k]);
            st.sum = __dp2a_lo(w[k], 0x0101u, st.sum);
            k1_px(w[k] & 0xffffu, c64, st.S[0], st.S[1], st.S[2]);
            k1_px(w[k] >> 16, c64, st.S[3], st.S[4], st.S[5]);
        }
    }
}

// ---- integer variant of the central sums -----------------------------------------------------------
// While |x - p| <= kK1IntLimit for every pixel of the tile, the three sums are exact integers that fit
// 64 bits per lane (up to 1,024 pixels per lane): y^2 in 32 bits (eight of them are added in 32 bits
// before they go into the 64-bit sum), y^3 by a signed and y^4 by an unsigned 32x32->64 multiply-add.
// That is 7 issue slots per pixel instead of 12 with FP64.  Whether the limit held is known only
// after the pass (from min and max); a tile that breaks it is done again with the FP64 loop.
constexpr int kK1IntLimit = 11585;                       // floor(2^13.5): 1,024 * y^4 < 2^64

struct K1IntState {
    uint32_t mn2, mx2, sum, cnt;
    unsigned long long S2;
    long long S3[2];
    unsigned long long S4[2];
};

__device__ __forceinline__ void k1_px_int(int y, uint32_t& s2, long long& S3, unsigned long long& S4) {
    const uint32_t y2 = (uint32_t)y * (uint32_t)y;      // exact below the limit, discarded above it
    s2 += y2;
    S3 += (long long)(int)y2 * (long long)y;             // IMAD.WIDE
    S4 += (unsigned long long)y2 * (unsigned long long)y2;   // IMAD.WIDE.U32
}

template <bool MASKED>
__device__ __forceinline__ void k1_vec_int(const uint4& v, const uint2& m, int p, K1IntState& st) {
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t h[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
    if (MASKED) {
        mask_halfwords(m.x, h[0], h[1]);
        mask_halfwords(m.y, h[2], h[3]);
        st.cnt += (__popc(h[0]) + __popc(h[1]) + __popc(h[2]) + __popc(h[3])) >> 4;
    }
    uint32_t s2 = 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (MASKED) {
            st.mn2 = __vminu2(st.mn2, w[k] | ~h[k]);
            w[k] &= h[k];
        } else {
            st.mn2 = __vminu2(st.mn2, w[k]);
        }
        st.mx2 = __vmaxu2(st.mx2, w[k]);
        st.sum = __dp2a_lo(w[k], 0x0101u, st.sum);
        int y0 = (int)(w[k] & 0xffffu) - p, y1 = (int)(w[k] >> 16) - p;
        if (MASKED) {                                      // outside the mask: y = 0 adds nothing
            y0 &= (int)__byte_perm(h[k], 0u, 0x1010);
            y1 &= (int)__byte_perm(h[k], 0u, 0x3232);
        }
        k1_px_int(y0, s2, st.S3[0], st.S4[0]);
        k1_px_int(y1, s2, st.S3[1], st.S4[1]);
    }
    st.S2 += s2;
}

// One tile's finished sums, parked until 32 of them can be turned into features by 32 lanes at once:
// the epilogue is ~250 instructions of FP64 divisions and square roots, a fifth of the per-tile cost
// when a single lane runs it.
struct K1Pending {
    double* o;                  // out_row + col_basic + 17 * slot
    uint32_t* status;
    uint32_t n_eff, vmin, vmax, total;
    long long p;
    double S2, S3, S4;
};

__device__ __forceinline__ void k1_epilogue(double* o, uint32_t* status, uint32_t n_eff,
                                            uint32_t vmin, uint32_t vmax, uint32_t total,
                                            long long p, double S2, double S3, double S4) {
    if (n_eff == 0) {
        const double nan = qnan();
        o[0] = nan; o[10] = nan; o[11] = nan; o[12] = nan; o[13] = nan; o[14] = nan; o[15] = nan;
        if (status) atomicOr(status, kStEmptyMask);
        return;
    }
    const double nn = (double)n_eff;
    const double mean = (double)total / nn;
    const double r = (double)((long long)total - p * (long long)n_eff);  // sum of y (exact)
    const double d = r / nn;                                             // mean - p
    // central sums from the pivoted sums: sum (y-d)^k
    const double C2 = S2 - r * d;
    const double C3 = S3 - 3.0 * d * S2 + 2.0 * r * d * d;
    const double C4 = S4 - 4.0 * d * S3 + 6.0 * d * d * S2 - 3.0 * r * d * d * d;
    const double m2 = C2 / nn, m3 = C3 / nn, m4 = C4 / nn;
    o[0] = (double)vmin;
    o[10] = (double)vmax;
    o[11] = (double)total;
    o[12] = mean;
    o[13] = sqrt(m2);
    const double thr = 2.220446049250313e-16 * mean;   // scipy: m2 <= (eps*mean)^2 -> NaN
    if (vmin == vmax || m2 <= thr * thr) {
        o[13] = (vmin == vmax) ? 0.0 : o[13];
        o[14] = qnan();
        o[15] = qnan();
        if (status) atomicOr(status, kStConstant);
    } else {
        o[14] = m4 / (m2 * m2) - 3.0;
        o[15] = m3 / (m2 * sqrt(m2));
    }
}

__device__ __forceinline__ void k1_epilogue(const Params& P, const Tile& T, uint32_t n_eff, uint32_t vmin,
                                            uint32_t vmax, uint32_t total, long long p, double S2, double S3, double S4) {
    k1_epilogue(T.out_row + P.col_basic + kNBasic * T.slot, T.status, n_eff, vmin, vmax, total, p, S2, S3, S4);
}
__device__ __forceinline__ void k1_park(K1Pending* slot, const Params& P, const Tile& T, uint32_t n_eff, uint32_t vmin,
                                        uint32_t vmax, uint32_t total, long long p, double S2, double S3, double S4) {
    slot->o = T.out_row + P.col_basic + kNBasic * T.slot; slot->status = T.status;
    slot->n_eff = n_eff; slot->vmin = vmin; slot->vmax = vmax; slot->total = total;
    slot->p = p; slot->S2 = S2; slot->S3 = S3; slot->S4 = S4;
}
__device__ __forceinline__ void k1_flush(const K1Pending* slots, int count, int lane) {
    __syncwarp();
    if (lane < count) {
        const K1Pending q = slots[lane];
        k1_epilogue(q.o, q.status, q.n_eff, q.vmin, q.vmax, q.total, q.p, q.S2, q.S3, q.S4);
    }
    __syncwarp();
}

template <bool MASKED>
__global__ void __launch_bounds__(256) k1_moments_kernel(const __grid_constant__ Params P) {
    const int lane = threadIdx.x & 31;
    __shared__ K1Pending pending_all[8][32];               // 256-thread CTAs: one row per warp
    K1Pending* pending = pending_all[threadIdx.x >> 5];
    int n_pending = 0;
    long long tnext = next_tile(P.sched + 0);
    while (tnext < P.n_tiles) {
        const long long t = tnext;
        tnext = next_tile(P.sched + 0);                    // one tile ahead
        const Tile T = resolve_tile(P, t);
        const uint4* px4 = reinterpret_cast<const uint4*>(T.px);
        const uint2* mk2 = reinterpret_cast<const uint2*>(T.mk);
        const int nfull = T.n >> 3, rem = T.n & 7;

        // ---- tail pixels (< 8) and the pivot sample: 32 vectors from the middle of the tile ----
        uint32_t xt = 0;
        bool tail_ok = false;
        if (lane < rem) {
            xt = T.px[nfull * 8 + lane];
            tail_ok = !MASKED || T.mk[nfull * 8 + lane] != 0;
        }
        uint32_t ssum = tail_ok ? xt : 0u, scnt = tail_ok ? 1u : 0u;
        uint32_t smn2 = 0xffffffffu, smx2 = 0u;            // extremes of the sample (packed 16-bit pairs)
        {
            // 32 vectors spread evenly over the tile: the pivot (their mean) and a first idea of the range
            // (stride + 1/2: a stride that is a multiple of the row length would sample one column only)
            const int stride = nfull >> 5;
            const int idx = nfull <= 32 ? lane : min((lane * (2 * stride + 1)) >> 1, nfull - 1);
            if (idx < nfull) {
                uint4 v = ld_reuse(px4 + idx);
                uint32_t h0 = 0xffffffffu, h1 = h0, h2 = h0, h3 = h0;
                if (MASKED) {
                    const uint2 m = __ldg(mk2 + idx);
                    mask_halfwords(m.x, h0, h1);
                    mask_halfwords(m.y, h2, h3);
                    scnt += (__popc(h0) + __popc(h1) + __popc(h2) + __popc(h3)) >> 4;
                } else {
                    scnt += 8;
                }
                smn2 = __vminu2(__vminu2(v.x | ~h0, v.y | ~h1), __vminu2(v.z | ~h2, v.w | ~h3));
                v.x &= h0; v.y &= h1; v.z &= h2; v.w &= h3;
                smx2 = __vmaxu2(__vmaxu2(v.x, v.y), __vmaxu2(v.z, v.w));
                ssum = __dp2a_lo(v.x, 0x0101u, ssum); ssum = __dp2a_lo(v.y, 0x0101u, ssum);
                ssum = __dp2a_lo(v.z, 0x0101u, ssum); ssum = __dp2a_lo(v.w, 0x0101u, ssum);
            }
        }
        const uint32_t smin = __reduce_min_sync(0xffffffffu, min(smn2 & 0xffffu, smn2 >> 16));
        const uint32_t smax = __reduce_max_sync(0xffffffffu, max(smx2 & 0xffffu, smx2 >> 16));
        ssum = __reduce_add_sync(0xffffffffu, ssum);
        scnt = __reduce_add_sync(0xffffffffu, scnt);
        if (MASKED && scnt == 0 && nfull > 0) {
            // the sample hit no masked pixel: take the pivot from a full pre-pass (re-read from L2)
            for (int idx = lane; idx < nfull; idx += 32) {
                uint4 v = ld_reuse(px4 + idx);
                const uint2 m = __ldg(mk2 + idx);
                uint32_t h0, h1, h2, h3;
                mask_halfwords(m.x, h0, h1);
                mask_halfwords(m.y, h2, h3);
                v.x &= h0; v.y &= h1; v.z &= h2; v.w &= h3;
                scnt += (__popc(h0) + __popc(h1) + __popc(h2) + __popc(h3)) >> 4;
                ssum = __dp2a_lo(v.x, 0x0101u, ssum); ssum = __dp2a_lo(v.y, 0x0101u, ssum);
                ssum = __dp2a_lo(v.z, 0x0101u, ssum); ssum = __dp2a_lo(v.w, 0x0101u, ssum);
            }
            ssum = __reduce_add_sync(0xffffffffu, ssum);
            scnt = __reduce_add_sync(0xffffffffu, scnt);
        }
        const long long p = scnt ? (long long)((ssum + (scnt >> 1)) / scnt) : 0;
        const unsigned long long c64 = (0x43300000ull << 32) | (unsigned long long)((1u << 20) - (uint32_t)p);

        // ---- the streaming pass: exact integer sums first (see K1IntState) ----
        bool done = false;
        // the sample already shows a range beyond the integer limit: go straight to the FP64 pass
        const bool sample_wide = scnt != 0u && smin <= smax &&
                                 ((int)smax - (int)p > kK1IntLimit || (int)p - (int)smin > kK1IntLimit);
        if (!P.k1_fp64_only && !sample_wide) {
            K1IntState st;
            st.mn2 = 0xffffffffu; st.mx2 = 0u; st.sum = 0u; st.cnt = 0u;
            st.S2 = 0ull; st.S3[0] = 0; st.S3[1] = 0; st.S4[0] = 0ull; st.S4[1] = 0ull;
            const int pi = (int)p;
            constexpr int kU = kK1Unroll;                  // loads in flight per lane (8 measured no faster)
            int idx = lane;
            for (; idx + 32 * (kU - 1) < nfull; idx += 32 * kU) {
                uint4 v[kU];
                uint2 m[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    v[u] = ld_stream(px4 + idx + 32 * u);
                    m[u] = make_uint2(0u, 0u);
                    if (MASKED) m[u] = __ldg(mk2 + idx + 32 * u);
                }
#pragma unroll
                for (int u = 0; u < kU; ++u) k1_vec_int<MASKED>(v[u], m[u], pi, st);
            }
            for (; idx < nfull; idx += 32) {
                const uint4 v = ld_stream(px4 + idx);
                uint2 m = make_uint2(0u, 0u);
                if (MASKED) m = __ldg(mk2 + idx);
                k1_vec_int<MASKED>(v, m, pi, st);
            }
            if (tail_ok) {
                st.mn2 = __vminu2(st.mn2, xt | 0xffff0000u);
                st.mx2 = __vmaxu2(st.mx2, xt);
                st.sum += xt;
                st.cnt += 1;
                uint32_t s2 = 0u;
                k1_px_int((int)xt - pi, s2, st.S3[0], st.S4[0]);
                st.S2 += s2;
            }
            const uint32_t total = __reduce_add_sync(0xffffffffu, st.sum);
            const uint32_t n_eff = MASKED ? __reduce_add_sync(0xffffffffu, st.cnt) : (uint32_t)T.n;
            const uint32_t vmin = __reduce_min_sync(0xffffffffu, min(st.mn2 & 0xffffu, st.mn2 >> 16));
            const uint32_t vmax = __reduce_max_sync(0xffffffffu, max(st.mx2 & 0xffffu, st.mx2 >> 16));
            // did every |x - p| stay within the limit?  (no pixel inside the mask: nothing was added)
            if (n_eff == 0 || ((int)vmax - pi <= kK1IntLimit && pi - (int)vmin <= kK1IntLimit)) {
                const unsigned long long S2 = warp_sum_redux(st.S2);
                const long long S3 = (long long)warp_sum_redux((unsigned long long)(st.S3[0] + st.S3[1]));
                // the lane sums fit 64 bits (see K1IntState), their total over the warp need not: combine in FP64
                const double S4 = warp_sum_redux_dbl(st.S4[0] + st.S4[1]);
                if (lane == 0) k1_park(pending + n_pending, P, T, n_eff, vmin, vmax, total, p, (double)S2, (double)S3, S4);
                if (++n_pending == 32) { k1_flush(pending, 32, lane); n_pending = 0; }
                done = true;
            }
        }
        if (done) continue;

        // ---- fallback: the same pass with FP64 sums (any 16-bit range) ----
        K1State st;
        st.mn2 = 0xffffffffu; st.mx2 = 0u; st.sum = 0u; st.cnt = 0u;
#pragma unroll
        for (int k = 0; k < 6; ++k) st.S[k] = 0.0;
        int idx = lane;
        for (; idx + 32 * (kK1Unroll - 1) < nfull; idx += 32 * kK1Unroll) {
            uint4 v[kK1Unroll];
            uint2 m[kK1Unroll];
#pragma unroll
            for (int u = 0; u < kK1Unroll; ++u) {
                v[u] = ld_stream(px4 + idx + 32 * u);
                m[u] = make_uint2(0u, 0u);
                if (MASKED) m[u] = __ldg(mk2 + idx + 32 * u);
            }
#pragma unroll
            for (int u = 0; u < kK1Unroll; ++u) k1_vec<MASKED>(v[u], m[u], c64, st);
        }
        for (; idx < nfull; idx += 32) {
            const uint4 v = ld_stream(px4 + idx);
            uint2 m = make_uint2(0u, 0u);
            if (MASKED) m = __ldg(mk2 + idx);
            k1_vec<MASKED>(v, m, c64, st);
        }
        if (tail_ok) {
            st.mn2 = __vminu2(st.mn2, xt | 0xffff0000u);
            st.mx2 = __vmaxu2(st.mx2, xt);
            st.sum += xt;
            st.cnt += 1;
            k1_px(xt, c64, st.S[0], st.S[1], st.S[2]);
        }

        // ---- warp reduction + epilogue ----
        const uint32_t total = __reduce_add_sync(0xffffffffu, st.sum);
        const uint32_t n_eff = MASKED ? __reduce_add_sync(0xffffffffu, st.cnt) : (uint32_t)T.n;
        const uint32_t vmin = __reduce_min_sync(0xffffffffu, min(st.mn2 & 0xffffu, st.mn2 >> 16));
        const uint32_t vmax = __reduce_max_sync(0xffffffffu, max(st.mx2 & 0xffffu, st.mx2 >> 16));
        const double S2 = warp_sum(st.S[0] + st.S[3]);
        const double S3 = warp_sum(st.S[1] + st.S[4]);
        const double S4 = warp_sum(st.S[2] + st.S[5]);
        if (lane == 0) k1_park(pending + n_pending, P, T, n_eff, vmin, vmax, total, p, S2, S3, S4);
        if (++n_pending == 32) { k1_flush(pending, 32, lane); n_pending = 0; }
    }
    k1_flush(pending, n_pending, lane);
}

}  // namespace imfeat

// -------------------------------------------------------------------------------------------------
// K1, bulk-copy variant (unmasked tiles, opt-in with IMFEAT_K1_TMA=1): a producer thread streams whole
// tiles into a shared-memory ring with cp.async.bulk (the 1-D TMA path, UBLKCP) completing on
// mbarriers; each consumer warp owns one tile at a time and reads it with conflict-free 128-bit
// shared loads.  Measured on B200 it is no faster than the direct-load kernel above (0.316 ms vs
// 0.296 ms per 10,000 objects): K1 is bound by instruction issue / the FP64 pipe, not by HBM
// latency, so the ring only removes stalls that occupancy already hides.
// -------------------------------------------------------------------------------------------------
namespace imfeat {

constexpr int kK1TmaConsumers = 16;                 // consumer warps per CTA
constexpr int kK1TmaThreads = 32 * (kK1TmaConsumers + 1);

__global__ void __launch_bounds__(kK1TmaThreads, 1)
k1_moments_tma_kernel(const __grid_constant__ Params P, int n_stages, int stage_bytes) {
    extern __shared__ __align__(128) unsigned char k1_smem[];
    // layout: [stages][full barriers][empty barriers]
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(k1_smem + (size_t)n_stages * stage_bytes);
    const uint32_t stage0 = smem_addr(k1_smem);
    const uint32_t full0 = smem_addr(bars), empty0 = smem_addr(bars + n_stages);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0)
        for (int s = 0; s < n_stages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const long long first = blockIdx.x;
    const long long mine = first < P.n_tiles ? (P.n_tiles - first + gridDim.x - 1) / gridDim.x : 0;

    if (warp == kK1TmaConsumers) {
        // ---- producer: one lane issues the bulk copies ----
        if (lane == 0) {
            TileWalk walk;
            walk.init(P, first < P.n_tiles ? first : 0, gridDim.x);
            for (long long k = 0; k < mine; ++k, walk.next()) {
                const int s = (int)(k % n_stages);
                const uint32_t ph = (uint32_t)((k / n_stages) & 1);
                mbar_wait(empty0 + 8 * s, ph ^ 1u);                       // slot free (first round passes)
                const Tile T = resolve_tile_rs(P, walk.row, walk.slot);
                const uint32_t bytes = ((uint32_t)T.n * 2u + 15u) & ~15u;  // padding inside the plane stride
                mbar_expect_tx(full0 + 8 * s, bytes);
                bulk_g2s(stage0 + (uint32_t)s * (uint32_t)stage_bytes, T.px, bytes, full0 + 8 * s);
            }
        }
        return;
    }

    // ---- consumers: warp w takes tiles w, w + CW, ... of this CTA ----
    TileWalk walk;
    walk.init(P, first + (long long)warp * gridDim.x < P.n_tiles ? first + (long long)warp * gridDim.x : 0,
              (long long)kK1TmaConsumers * gridDim.x);
    for (long long k = warp; k < mine; k += kK1TmaConsumers, walk.next()) {
        const int s = (int)(k % n_stages);
        const uint32_t ph = (uint32_t)((k / n_stages) & 1);
        const Tile T = resolve_tile_rs(P, walk.row, walk.slot);
        const int nfull = T.n >> 3, rem = T.n & 7;
        const uint4* px4 = reinterpret_cast<const uint4*>(k1_smem + (size_t)s * stage_bytes);
        const uint16_t* px1 = reinterpret_cast<const uint16_t*>(px4);
        mbar_wait(full0 + 8 * s, ph);                                     // tile landed

        uint32_t xt = 0;
        const bool tail_ok = lane < rem;
        if (tail_ok) xt = px1[nfull * 8 + lane];
        uint32_t ssum = tail_ok ? xt : 0u, scnt = tail_ok ? 1u : 0u;
        {
            const int s0 = nfull > 32 ? (nfull >> 1) - 16 : 0;
            const int idx = s0 + lane;
            if (idx < nfull) {
                const uint4 v = px4[idx];
                scnt += 8;
                ssum = __dp2a_lo(v.x, 0x0101u, ssum); ssum = __dp2a_lo(v.y, 0x0101u, ssum);
                ssum = __dp2a_lo(v.z, 0x0101u, ssum); ssum = __dp2a_lo(v.w, 0x0101u, ssum);
            }
        }
        ssum = __reduce_add_sync(0xffffffffu, ssum);
        scnt = __reduce_add_sync(0xffffffffu, scnt);
        const long long p = scnt ? (long long)((ssum + (scnt >> 1)) / scnt) : 0;
        const unsigned long long c64 = (0x43300000ull << 32) | (unsigned long long)((1u << 20) - (uint32_t)p);

        K1State st;
        st.mn2 = 0xffffffffu; st.mx2 = 0u; st.sum = 0u; st.cnt = 0u;
#pragma unroll
        for (int q = 0; q < 6; ++q) st.S[q] = 0.0;
        const uint2 nomask = make_uint2(0u, 0u);
        int idx = lane;
        for (; idx + 32 < nfull; idx += 64) {
            const uint4 v0 = px4[idx], v1 = px4[idx + 32];
            k1_vec<false>(v0, nomask, c64, st);
            k1_vec<false>(v1, nomask, c64, st);
        }
        for (; idx < nfull; idx += 32) k1_vec<false>(px4[idx], nomask, c64, st);
        if (tail_ok) {
            st.mn2 = __vminu2(st.mn2, xt | 0xffff0000u);
            st.mx2 = __vmaxu2(st.mx2, xt);
            st.sum += xt;
            k1_px(xt, c64, st.S[0], st.S[1], st.S[2]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * s);                       // stage can be refilled

        const uint32_t total = __reduce_add_sync(0xffffffffu, st.sum);
        const uint32_t vmin = __reduce_min_sync(0xffffffffu, min(st.mn2 & 0xffffu, st.mn2 >> 16));
        const uint32_t vmax = __reduce_max_sync(0xffffffffu, max(st.mx2 & 0xffffu, st.mx2 >> 16));
        const double S2 = warp_sum(st.S[0] + st.S[3]);
        const double S3 = warp_sum(st.S[1] + st.S[4]);
        const double S4 = warp_sum(st.S[2] + st.S[5]);
        if (lane == 0) k1_epilogue(P, T, (uint32_t)T.n, vmin, vmax, total, p, S2, S3, S4);
    }
}

}  // namespace imfeat
