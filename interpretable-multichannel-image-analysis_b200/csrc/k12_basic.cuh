// k12_basic.cuh -- K12: the whole basic block (17 columns) in ONE pass over the tile, one WARP per tile.
//
// Replaces, per channel (NB = reference notebook raw line):
//   min/max NB:241,251  nine np.percentile NB:242-250  total NB:254  mean NB:257  std NB:258
//   kurtosis NB:259  skew NB:260  shannon_entropy NB:262
//
// K1 (k1_moments.cuh) and K2c (k2_order_entropy.cuh) each streamed the tile once, K2c after K1 because its
// 4,096-bin histogram is relative to the tile minimum.  Here the histogram window is fixed BEFORE the pass,
// from the same 256-pixel sample that gives K1 its pivot: the window of 4,096 values is centred on the
// sample's range, so a tile whose true range sticks out of the sample's by less than the slack on either
// side -- every tile of 12-bit data -- needs no second look at its pixels.  Whether the window held is known
// exactly after the pass (from the minimum and maximum K1 computes anyway); a tile that broke it has its
// 8 KB histogram wiped and goes to the worklist of the full-range kernel, like K2c's wide tiles before.
// Per pixel the pass costs K1's integer sums (7 issue slots) + one fire-and-forget shared-memory atomic with
// its address arithmetic (5-6), instead of two passes with their own loads, mask expansion and loop.
#pragma once
#include "k1_moments.cuh"
#include "k2_order_entropy.cuh"

namespace imfeat {

constexpr int kK12Clc = 64;     // counts below this take c * log2(c) from shared memory
struct K12Smem {
    K2cSmem h;                  // 4,096-bin histogram (16-bit counters) + percentile scratch of this warp
#ifndef IMFEAT_K12_PARK
#define IMFEAT_K12_PARK 16               // 16 instead of 32: 1.4 KB less shared memory per warp, 21 instead of 19 warps per SM (K12 0.581 -> 0.573 ms)
#endif
    K1Pending pending[IMFEAT_K12_PARK];      // finished tiles whose moment epilogues run side by side
    double clc[kK12Clc];        // c * log2(c), c < kK12Clc (0 for c = 0, 1)
};

// One pixel into the warp's histogram (fire-and-forget).  inb = all ones when the pixel counts (inside the mask), else 0:
// then the increment is 0, on word `lane` (distinct banks; equal background values outside the mask would
// otherwise serialise on one address).  Two issue slots fewer than k2c_px: the increment 1 << 16 * (bin & 1) is one
// wrapping funnel shift (only the low five bits of the amount count), with the mask bit as its source.
template <bool MASKED>
__device__ __forceinline__ void k12_px(K2cSmem& S, uint32_t x, uint32_t inb, uint32_t base, uint32_t lane4) {
    const uint32_t bin = x - base;
    uint32_t off = (bin << 1) & (uint32_t)(kK2cWords * 4 - 4);
    const uint32_t inc = __funnelshift_l(0u, MASKED ? (inb & 1u) : 1u, bin << 4);
    if (MASKED) off = (off & inb) | (lane4 & ~inb);
#ifdef IMFEAT_CHECKS_SELFTEST
    IMFEAT_CHECK(off < 64u);                               // control: must fire (shows that the checks are live)
#endif
    IMFEAT_CHECK(off < (uint32_t)(kK2cWords * 4) && (off & 3u) == 0u);
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(smem_addr(S.hist) + off), "r"(inc) : "memory");
}

// the eight pixels of one 128-bit load: K1's exact integer sums, and (HIST) one histogram add per pixel
template <bool MASKED, bool HIST>
__device__ __forceinline__ void k12_vec(K2cSmem& S, const uint4& v, const uint2& m, int p, uint32_t base, K1IntState& st) {
    // eight pixels outside the mask add nothing to any sum, to the extrema or to the histogram: skip them.  Rows
    // above and below a mask's bounding box are whole warp iterations (32 lanes x 8 pixels = 4 rows of 64).
    if (MASKED && (m.x | m.y) == 0u) return;
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t h[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
    uint32_t nz[2] = {0xffffffffu, 0xffffffffu};
    if (MASKED) {
        nz[0] = __vcmpne4(m.x, 0u); nz[1] = __vcmpne4(m.y, 0u);          // 0xff per pixel inside the mask
        h[0] = __byte_perm(nz[0], 0u, 0x1100); h[1] = __byte_perm(nz[0], 0u, 0x3322);
        h[2] = __byte_perm(nz[1], 0u, 0x1100); h[3] = __byte_perm(nz[1], 0u, 0x3322);
        st.cnt += (__popc(nz[0]) + __popc(nz[1])) >> 3;
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
        const uint32_t x0 = w[k] & 0xffffu, x1 = w[k] >> 16;
        int y0 = (int)x0 - p, y1 = (int)x1 - p;
        if (MASKED) {                                      // outside the mask: y = 0 adds nothing
            y0 &= (int)__byte_perm(h[k], 0u, 0x1010);
            y1 &= (int)__byte_perm(h[k], 0u, 0x3232);
        }
        k1_px_int(y0, s2, st.S3[0], st.S4[0]);
        k1_px_int(y1, s2, st.S3[1], st.S4[1]);
        if (HIST) {
            const uint32_t lane4 = 4u * (threadIdx.x & 31);
            k12_px<MASKED>(S, x0, __byte_perm(h[k], 0u, 0x1010), base, lane4);
            k12_px<MASKED>(S, x1, __byte_perm(h[k], 0u, 0x3232), base, lane4);
        }
    }
    st.S2 += s2;
}

// Percentiles (numpy "linear", bit for bit) and entropy from the warp's histogram whose bin 0 is the value
// `base`; the values present lie in [vmin, vmax].  Clears the used range.  (K2c's second half, with the walk
// starting at the block of the minimum instead of bin 0.)
__device__ __forceinline__ void k12_order_entropy(K2cSmem& S, const double* clc, const Params& P, double* o, int n, uint32_t base,
                                                  uint32_t vmin, uint32_t vmax, int lane) {
    const int b_lo = (int)(vmin - base), b_hi = (int)(vmax - base);
    {
        int lo[9], hi[9], maxrank = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const double virt = __dmul_rn((double)(n - 1), P.quant[k]);
            if (virt >= (double)(n - 1)) { lo[k] = hi[k] = n - 1; }
            else { lo[k] = (int)floor(virt); hi[k] = lo[k] + 1; }
            maxrank = max(maxrank, hi[k]);
        }
        int cum = 0;
        for (int block = b_lo >> 6; block * 64 <= b_hi; ++block) {
            const uint32_t wv = S.hist[block * 32 + lane];
            const int c0 = wv & 0xffffu, c1 = wv >> 16, tot = c0 + c1;
            int incl = tot;
#pragma unroll
            for (int o2 = 1; o2 < 32; o2 <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o2);
                if (lane >= o2) incl += v;
            }
            const int r0 = cum + incl - tot, r1 = cum + incl;
            const int bval = (int)base + block * 64 + 2 * lane;
            if (tot != 0 && r0 <= maxrank) {               // near the minimum most words are empty
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    if (lo[k] >= r0 && lo[k] < r1) S.vals[2 * k] = bval + (lo[k] >= r0 + c0);
                    if (hi[k] >= r0 && hi[k] < r1) S.vals[2 * k + 1] = bval + (hi[k] >= r0 + c0);
                }
            }
            cum += __shfl_sync(0xffffffffu, incl, 31);
            if (cum > maxrank) break;
        }
        __syncwarp();
        if (lane < 9) {
            const double virt = __dmul_rn((double)(n - 1), P.quant[lane]);
            const double g = virt - floor(virt);
            const int a = S.vals[2 * lane], b = S.vals[2 * lane + 1];
            const double diff = (double)(b - a);
            // numpy _lerp: a + diff*t, replaced by b - diff*(1-t) where t >= 0.5 (no FMA there)
            o[1 + lane] = (g >= 0.5) ? __dsub_rn((double)b, __dmul_rn(diff, __dsub_rn(1.0, g)))
                                     : __dadd_rn((double)a, __dmul_rn(diff, g));
        }
    }
    __syncwarp();
    // entropy = log2 n - (1/n) sum_bins c log2 c, in the pass that clears the used range: fixed order per lane
    // and a fixed shuffle tree, so the double sum is reproducible bit for bit
    // (four words per lane and step; a count of 0 or 1 adds nothing -- most bins of a tile hold at most one pixel)
    double hs = 0.0;
    uint4* hist4 = reinterpret_cast<uint4*>(S.hist);
    for (int k = (b_lo >> 3) + lane; k <= (b_hi >> 3); k += 32) {
        const uint4 q = hist4[k];
        if ((q.x | q.y | q.z | q.w) == 0u) continue;
        hist4[k] = make_uint4(0u, 0u, 0u, 0u);
        const uint32_t wv4[4] = {q.x, q.y, q.z, q.w};
        // small counts (all but flat backgrounds): c * log2(c) from shared memory, no branch per count
        if (((q.x | q.y | q.z | q.w) & ~(0x00010001u * (uint32_t)(kK12Clc - 1))) == 0u) {
#pragma unroll
            for (int u = 0; u < 4; ++u) hs += clc[wv4[u] & 0xffffu] + clc[wv4[u] >> 16];
            continue;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t c0 = wv4[u] & 0xffffu, c1 = wv4[u] >> 16;
            if (c0 > 1u) hs = fma((double)c0, __ldg(P.log2tab + c0), hs);
            if (c1 > 1u) hs = fma((double)c1, __ldg(P.log2tab + c1), hs);
        }
    }
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) hs += __shfl_xor_sync(0xffffffffu, hs, o2);
    if (lane == 0) o[16] = vmin == vmax ? 0.0 : __ldg(P.log2tab + n) - hs / (double)n;   // one value only: entropy is exactly 0
    __syncwarp();
}

#ifndef IMFEAT_K12_WARPS
#define IMFEAT_K12_WARPS 20            // resident warps per SM the register budget is set for
#endif
template <bool MASKED>
__global__ void __launch_bounds__(32, IMFEAT_K12_WARPS) k12_basic_kernel(const __grid_constant__ Params P,
                                                           uint32_t* __restrict__ worklist,
                                                           uint32_t* __restrict__ worklist_count) {
    __shared__ K12Smem SS;
    K2cSmem& S = SS.h;
    K1Pending* pending = SS.pending;
    const int lane = threadIdx.x;
    for (int k = lane; k < kK2cWords; k += 32) S.hist[k] = 0u;
    for (int k = lane; k < kK12Clc; k += 32) SS.clc[k] = (double)k * __ldg(P.log2tab + k);      // log2tab[0] = 0
    __syncwarp();
    int n_pending = 0;
    long long tnext = next_tile(P.sched + 0);
    while (tnext < P.n_tiles) {
        const long long t = tnext;
        tnext = next_tile(P.sched + 0);                    // one tile ahead
        const Tile T = resolve_tile(P, t);
        double* o = T.out_row + P.col_basic + kNBasic * T.slot;
        const uint4* px4 = reinterpret_cast<const uint4*>(T.px);
        const uint2* mk2 = reinterpret_cast<const uint2*>(T.mk);
        const int nfull = T.n >> 3, rem = T.n & 7;

        // ---- tail pixels (< 8) and the sample: 32 vectors spread over the tile give the pivot and the window ----
        uint32_t xt = 0;
        bool tail_ok = false;
        if (lane < rem) {
            xt = T.px[nfull * 8 + lane];
            tail_ok = !MASKED || T.mk[nfull * 8 + lane] != 0;
        }
        uint32_t ssum = tail_ok ? xt : 0u, scnt = tail_ok ? 1u : 0u;
        uint32_t smn2 = tail_ok ? (xt | 0xffff0000u) : 0xffffffffu, smx2 = tail_ok ? xt : 0u;
        {
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
                smn2 = __vminu2(smn2, __vminu2(__vminu2(v.x | ~h0, v.y | ~h1), __vminu2(v.z | ~h2, v.w | ~h3)));
                v.x &= h0; v.y &= h1; v.z &= h2; v.w &= h3;
                smx2 = __vmaxu2(smx2, __vmaxu2(__vmaxu2(v.x, v.y), __vmaxu2(v.z, v.w)));
                ssum = __dp2a_lo(v.x, 0x0101u, ssum); ssum = __dp2a_lo(v.y, 0x0101u, ssum);
                ssum = __dp2a_lo(v.z, 0x0101u, ssum); ssum = __dp2a_lo(v.w, 0x0101u, ssum);
            }
        }
        ssum = __reduce_add_sync(0xffffffffu, ssum);
        scnt = __reduce_add_sync(0xffffffffu, scnt);
        if (MASKED && scnt == 0 && nfull > 0) {
            // the sample hit no masked pixel: pivot and range from a full pre-pass (re-read from L2)
            for (int idx = lane; idx < nfull; idx += 32) {
                uint4 v = ld_reuse(px4 + idx);
                const uint2 m = __ldg(mk2 + idx);
                uint32_t h0, h1, h2, h3;
                mask_halfwords(m.x, h0, h1);
                mask_halfwords(m.y, h2, h3);
                smn2 = __vminu2(smn2, __vminu2(__vminu2(v.x | ~h0, v.y | ~h1), __vminu2(v.z | ~h2, v.w | ~h3)));
                v.x &= h0; v.y &= h1; v.z &= h2; v.w &= h3;
                smx2 = __vmaxu2(smx2, __vmaxu2(__vmaxu2(v.x, v.y), __vmaxu2(v.z, v.w)));
                scnt += (__popc(h0) + __popc(h1) + __popc(h2) + __popc(h3)) >> 4;
                ssum = __dp2a_lo(v.x, 0x0101u, ssum); ssum = __dp2a_lo(v.y, 0x0101u, ssum);
                ssum = __dp2a_lo(v.z, 0x0101u, ssum); ssum = __dp2a_lo(v.w, 0x0101u, ssum);
            }
            ssum = __reduce_add_sync(0xffffffffu, ssum);
            scnt = __reduce_add_sync(0xffffffffu, scnt);
        }
        const uint32_t smin = __reduce_min_sync(0xffffffffu, min(smn2 & 0xffffu, smn2 >> 16));
        const uint32_t smax = __reduce_max_sync(0xffffffffu, max(smx2 & 0xffffu, smx2 >> 16));
        const long long p = scnt ? (long long)((ssum + (scnt >> 1)) / scnt) : 0;
        const int pi = (int)p;
        const bool have = scnt != 0u && smin <= smax;
        // the sample already shows a range beyond the integer limit: go straight to the FP64 pass
        const bool sample_wide = have && ((int)smax - pi > kK1IntLimit || pi - (int)smin > kK1IntLimit);
        // histogram window [base, base + 4096): centred on the sample's range
        const int slack = kK2cBins - ((int)smax - (int)smin + 1);
        const bool hist_on = have && slack >= 0 && !sample_wide && !P.k1_fp64_only;
        const uint32_t base = hist_on ? (uint32_t)min(max((int)smin - (slack >> 1), 0), 65536 - kK2cBins) : 0u;

        bool moments_done = false, order_done = false;
        uint32_t n_eff = 0u, vmin = 0u, vmax = 0u;
        if (!P.k1_fp64_only && !sample_wide) {
            // ---- the pass: exact integer central sums (see K1IntState) and, when the window is on, the histogram ----
            K1IntState st;
            st.mn2 = 0xffffffffu; st.mx2 = 0u; st.sum = 0u; st.cnt = 0u;
            st.S2 = 0ull; st.S3[0] = 0; st.S3[1] = 0; st.S4[0] = 0ull; st.S4[1] = 0ull;
            constexpr int kU = kK1Unroll;                  // loads in flight per lane
            int idx = lane;
            if (hist_on) {
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
                    for (int u = 0; u < kU; ++u) k12_vec<MASKED, true>(S, v[u], m[u], pi, base, st);
                }
                for (; idx < nfull; idx += 32) {
                    const uint4 v = ld_stream(px4 + idx);
                    uint2 m = make_uint2(0u, 0u);
                    if (MASKED) m = __ldg(mk2 + idx);
                    k12_vec<MASKED, true>(S, v, m, pi, base, st);
                }
                if (tail_ok) k12_px<false>(S, xt, 0xffffffffu, base, 0u);
            } else {
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
                    for (int u = 0; u < kU; ++u) k12_vec<MASKED, false>(S, v[u], m[u], pi, base, st);
                }
                for (; idx < nfull; idx += 32) {
                    const uint4 v = ld_stream(px4 + idx);
                    uint2 m = make_uint2(0u, 0u);
                    if (MASKED) m = __ldg(mk2 + idx);
                    k12_vec<MASKED, false>(S, v, m, pi, base, st);
                }
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
            n_eff = MASKED ? __reduce_add_sync(0xffffffffu, st.cnt) : (uint32_t)T.n;
            vmin = __reduce_min_sync(0xffffffffu, min(st.mn2 & 0xffffu, st.mn2 >> 16));
            vmax = __reduce_max_sync(0xffffffffu, max(st.mx2 & 0xffffu, st.mx2 >> 16));
            // did every |x - p| stay within the limit?  (no pixel inside the mask: nothing was added)
            if (n_eff == 0 || ((int)vmax - pi <= kK1IntLimit && pi - (int)vmin <= kK1IntLimit)) {
                const unsigned long long S2 = warp_sum_redux(st.S2);
                const long long S3 = (long long)warp_sum_redux((unsigned long long)(st.S3[0] + st.S3[1]));
                // the lane sums fit 64 bits (see K1IntState), their total over the warp need not: combine in FP64
                const double S4 = warp_sum_redux_dbl(st.S4[0] + st.S4[1]);
                if (lane == 0) k1_park(pending + n_pending, P, T, n_eff, vmin, vmax, total, p, (double)S2, (double)S3, S4);
                ++n_pending;
                moments_done = true;
            }
            __syncwarp();                                  // the histogram adds of all lanes are done
            if (hist_on) {
                if (n_eff != 0u && vmin >= base && vmax < base + (uint32_t)kK2cBins) {
                    k12_order_entropy(S, SS.clc, P, o, (int)n_eff, base, vmin, vmax, lane);
                    order_done = true;
                } else if (n_eff != 0u) {
                    // the window did not hold: the counts are meaningless (offsets wrapped around) -- wipe them
                    for (int k = lane; k < kK2cWords; k += 32) S.hist[k] = 0u;
                    __syncwarp();
                }
            }
        }
        if (!moments_done) {
            // ---- fallback: the same pass with FP64 sums (any 16-bit range) ----
            const unsigned long long c64 = (0x43300000ull << 32) | (unsigned long long)((1u << 20) - (uint32_t)p);
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
                    v[u] = ld_reuse(px4 + idx + 32 * u);
                    m[u] = make_uint2(0u, 0u);
                    if (MASKED) m[u] = __ldg(mk2 + idx + 32 * u);
                }
#pragma unroll
                for (int u = 0; u < kK1Unroll; ++u) k1_vec<MASKED>(v[u], m[u], c64, st);
            }
            for (; idx < nfull; idx += 32) {
                const uint4 v = ld_reuse(px4 + idx);
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
            const uint32_t total = __reduce_add_sync(0xffffffffu, st.sum);
            n_eff = MASKED ? __reduce_add_sync(0xffffffffu, st.cnt) : (uint32_t)T.n;
            vmin = __reduce_min_sync(0xffffffffu, min(st.mn2 & 0xffffu, st.mn2 >> 16));
            vmax = __reduce_max_sync(0xffffffffu, max(st.mx2 & 0xffffu, st.mx2 >> 16));
            const double S2 = warp_sum(st.S[0] + st.S[3]);
            const double S3 = warp_sum(st.S[1] + st.S[4]);
            const double S4 = warp_sum(st.S[2] + st.S[5]);
            if (lane == 0) k1_park(pending + n_pending, P, T, n_eff, vmin, vmax, total, p, S2, S3, S4);
            ++n_pending;
        }
        if (!order_done) {
            if (n_eff == 0u) {                             // no pixel inside the mask
                if (lane == 0) {
                    const double nan = qnan();
#pragma unroll
                    for (int q = 1; q <= 9; ++q) o[q] = nan;
                    o[16] = nan;
                }
            } else if (lane == 0) {                        // left to the full-range kernel
                worklist[atomicAdd(worklist_count, 1u)] = (uint32_t)t;
            }
        }
        if (n_pending == IMFEAT_K12_PARK) { k1_flush(pending, IMFEAT_K12_PARK, lane); n_pending = 0; }
    }
    k1_flush(pending, n_pending, lane);
}

}  // namespace imfeat
