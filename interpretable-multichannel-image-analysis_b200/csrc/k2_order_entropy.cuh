// k2_order_entropy.cuh -- K2: order statistics + Shannon entropy, one CTA per tile.
//
// Replaces, per channel:  np.percentile(X, 0.1 .. 0.9)  NB:242-250   shannon_entropy(X)  NB:262
//
// Both need the multiplicity of every raw 16-bit value in the tile.  The CTA keeps a
// privatised 65,536-bin histogram in shared memory (16-bit counters, two per word, 128 KB),
// built with shared-memory atomics, plus a 1,024-bit "occupied 64-value block" bitmap.
//   entropy      H = (1/n) * sum_pixels (log2 n - log2 c[x_p])        (no pass over the bins)
//   percentiles  one warp walks only the occupied blocks from the bottom until the needed
//                ranks are covered, then applies numpy's linear-interpolation formula
//   clear        by re-walking the pixels, never densely
#pragma once
#include "common.cuh"

namespace imfeat {

constexpr int kK2Threads = 512;
constexpr int kK2Warps = kK2Threads / 32;

struct K2Smem {
    uint32_t hist[32768];
    uint32_t coarse[32];
    int vals[18];
    uint32_t wcnt[kK2Warps];
    double wpart[kK2Warps];
};

__device__ __forceinline__ void k2_add(K2Smem& S, uint32_t x) {
    atomicAdd(&S.hist[x >> 1], 1u << ((x & 1u) << 4));
    const uint32_t b = x >> 6, bit = 1u << (b & 31u);
    if (!(*(volatile uint32_t*)&S.coarse[b >> 5] & bit)) atomicOr(&S.coarse[b >> 5], bit);
}
__device__ __forceinline__ uint32_t k2_count(const K2Smem& S, uint32_t x) {
    return (S.hist[x >> 1] >> ((x & 1u) << 4)) & 0xffffu;
}

// PHASE 0: histogram build, 1: entropy read-back, 2: sparse clear.
template <int PHASE, bool MASKED>
__device__ __forceinline__ void k2_px(K2Smem& S, const Params& P, uint32_t x, bool ok,
                                      uint32_t& cnt, double& acc, double log2n) {
    if (MASKED && !ok) return;
    if (PHASE == 0) { k2_add(S, x); if (MASKED) ++cnt; }
    if (PHASE == 1) acc += log2n - __ldg(P.log2tab + k2_count(S, x));
    if (PHASE == 2) S.hist[x >> 1] = 0u;
}

template <int PHASE, bool MASKED>
__device__ __forceinline__ void k2_walk(K2Smem& S, const Params& P, const Tile& T, uint32_t& cnt,
                                        double& acc, double log2n) {
    const uint4* px4 = reinterpret_cast<const uint4*>(T.px);
    const uint2* mk2 = reinterpret_cast<const uint2*>(T.mk);
    const int nfull = T.n >> 3, rem = T.n & 7;
    for (int idx = threadIdx.x; idx < nfull; idx += kK2Threads) {
        const uint4 v = ld_reuse(px4 + idx);
        uint2 m = make_uint2(0x01010101u, 0x01010101u);
        if (MASKED) m = __ldg(mk2 + idx);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t mb = (k < 2 ? m.x : m.y) >> (16 * (k & 1));
            k2_px<PHASE, MASKED>(S, P, w[k] & 0xffffu, (mb & 0xffu) != 0, cnt, acc, log2n);
            k2_px<PHASE, MASKED>(S, P, w[k] >> 16, (mb & 0xff00u) != 0, cnt, acc, log2n);
        }
    }
    if ((int)threadIdx.x < rem) {
        const int i = nfull * 8 + threadIdx.x;
        k2_px<PHASE, MASKED>(S, P, T.px[i], !MASKED || T.mk[i] != 0, cnt, acc, log2n);
    }
}

// Warp 0: numpy percentile (method "linear") from the histogram; ranks are 0-based positions
// in the sorted multiset.
__device__ __forceinline__ void k2_percentiles(K2Smem& S, const Params& P, int n, double* o) {
    const int lane = threadIdx.x & 31;
    int lo[9], hi[9], maxrank = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const double virt = __dmul_rn((double)(n - 1), P.quant[k]);
        if (virt >= (double)(n - 1)) { lo[k] = hi[k] = n - 1; }
        else { lo[k] = (int)floor(virt); hi[k] = lo[k] + 1; }
        maxrank = max(maxrank, hi[k]);
    }
    int cum = 0;
    bool done = false;
    for (int cw = 0; cw < 32 && !done; ++cw) {
        uint32_t bits = S.coarse[cw];
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int block = cw * 32 + b;
            const uint32_t wv = S.hist[block * 32 + lane];
            const int c0 = wv & 0xffffu, c1 = wv >> 16, tot = c0 + c1;
            int incl = tot;
#pragma unroll
            for (int o2 = 1; o2 < 32; o2 <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o2);
                if (lane >= o2) incl += v;
            }
            const int r0 = cum + incl - tot, r1 = cum + incl;
            const int base = block * 64 + 2 * lane;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                if (lo[k] >= r0 && lo[k] < r1) S.vals[2 * k] = base + (lo[k] >= r0 + c0);
                if (hi[k] >= r0 && hi[k] < r1) S.vals[2 * k + 1] = base + (hi[k] >= r0 + c0);
            }
            cum += __shfl_sync(0xffffffffu, incl, 31);
            if (cum > maxrank) { done = true; break; }
        }
    }
    __syncwarp();
    if (lane < 9) {
        const double virt = __dmul_rn((double)(n - 1), P.quant[lane]);
        const double g = virt - floor(virt);
        const int a = S.vals[2 * lane], b = S.vals[2 * lane + 1];
        const double diff = (double)(b - a);
        // numpy _lerp: a + diff*t, replaced by b - diff*(1-t) where t >= 0.5 (no FMA there)
        const double r = (g >= 0.5) ? __dsub_rn((double)b, __dmul_rn(diff, __dsub_rn(1.0, g)))
                                    : __dadd_rn((double)a, __dmul_rn(diff, g));
        o[1 + lane] = r;
    }
}

template <bool MASKED>
__global__ void __launch_bounds__(kK2Threads, 1) k2_order_entropy_kernel(const __grid_constant__ Params P) {
    extern __shared__ __align__(16) unsigned char k2_smem_raw[];
    K2Smem& S = *reinterpret_cast<K2Smem*>(k2_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int k = tid; k < 32768; k += kK2Threads) S.hist[k] = 0u;
    if (tid < 32) S.coarse[tid] = 0u;
    __syncthreads();

    for (long long t = blockIdx.x; t < P.n_tiles; t += gridDim.x) {
        const Tile T = resolve_tile(P, t);
        double* o = T.out_row + P.col_basic + kNBasic * T.slot;
        uint32_t cnt = 0;
        double acc = 0.0;

        k2_walk<0, MASKED>(S, P, T, cnt, acc, 0.0);
        if (MASKED) {
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (lane == 0) S.wcnt[warp] = cnt;
        }
        __syncthreads();                                   // histogram complete
        int n = T.n;
        if (MASKED) {
            n = 0;
#pragma unroll
            for (int k = 0; k < kK2Warps; ++k) n += (int)S.wcnt[k];
        }
        if (n > 0) {
            if (warp == 0) k2_percentiles(S, P, n, o);
            const double log2n = __ldg(P.log2tab + n);
            k2_walk<1, MASKED>(S, P, T, cnt, acc, log2n);
        }
        acc = warp_sum(acc);
        if (lane == 0) S.wpart[warp] = acc;
        __syncthreads();                                   // all read-backs done
        k2_walk<2, MASKED>(S, P, T, cnt, acc, 0.0);
        if (tid < 32) S.coarse[tid] = 0u;
        if (tid == 0) {
            if (n > 0) {
                double tot = 0.0;
#pragma unroll
                for (int k = 0; k < kK2Warps; ++k) tot += S.wpart[k];
                o[16] = tot / (double)n;
            } else {
                const double nan = qnan();
#pragma unroll
                for (int k = 1; k <= 9; ++k) o[k] = nan;
                o[16] = nan;
            }
        }
        __syncthreads();                                   // table clean for the next tile
    }
}

}  // namespace imfeat
