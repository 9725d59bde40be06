// k2b_bytes.cuh -- K2b: order statistics + Shannon entropy of full-range tiles with PRIVATE byte-counter tables.
//
// Replaces, per channel:  np.percentile(X, q1 .. q9)  NB:242-250   shannon_entropy(X)  NB:262
// for the tiles whose values left K12's 4,096-value window (genuine 16-bit data).
//
// K2 (k2_order_entropy.cuh) keeps ONE 65,536-bin table of 16-bit counters per SM (128 KB) that its thread groups
// take in turns; once everything else had been moved out of the table phase (round 2), the phase itself was what
// was left: 35 % of all stall samples were groups waiting for the table.  A tile of a few thousand pixels almost
// never holds one value more than 255 times, so here a CTA of 256 threads owns a 64 KB table of 8-BIT counters
// (four bins per word) and three CTAs share an SM with no hand-over at all.  The returning atomic that builds the
// histogram tells every pixel its bin's count before it: the entropy terms G[old] come from it as in K2, and an old
// count of 255 is the exact sign that this bin wraps -- the tile is then left, untouched in the output, to K2 itself
// through a second worklist (flat regions of saturated pixels are where that happens).
//   percentiles  two levels as in K2: the 256 counts by high byte next to the table, then one 256-byte slice of
//                the table per rank; every warp takes the percentiles k = warp, warp + 8
//   clear        by re-walking the pixels (word stores: a wrapped counter carries into its neighbour), never densely
#pragma once
#include "k2_order_entropy.cuh"

namespace imfeat {

constexpr int kK2bThreads = 256;
constexpr int kK2bWarps = kK2bThreads / 32;
constexpr int kK2bVec = 2;          // 16-byte vectors per thread kept in registers (a 64x64 tile exactly)

struct alignas(16) K2bSmem {
    uint32_t hist[16384];           // 65,536 byte counters: value x -> byte x & 3 of word x >> 2
    uint32_t c256[256];             // pixels by value >> 8
    unsigned long long gfix[256];   // G[k] = (k+1) log2(k+1) - k log2 k, 2^-42 fixed point, k < 256
    unsigned long long wacc[kK2bWarps];
    uint32_t cnt, maxold, next[2];  // next: the CTA's next worklist position, by iteration parity
};

// one pixel: count of its value before this increment (0 for a pixel outside the mask, which adds 0 to its own bin)
__device__ __forceinline__ uint32_t k2b_add(K2bSmem& S, uint32_t x, bool in) {
    const uint32_t sh = (x & 3u) << 3;
    const uint32_t old = atomicAdd(&S.hist[x >> 2], in ? (1u << sh) : 0u);
    atomicAdd(&S.c256[x >> 8], in ? 1u : 0u);
    return in ? ((old >> sh) & 0xffu) : 0u;
}

template <int PHASE, bool MASKED>
__device__ __forceinline__ void k2b_vec(K2bSmem& S, const uint4& v, const uint2& m, uint32_t& cnt, uint32_t& maxold,
                                        unsigned long long& acc) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t mb = (k < 2 ? m.x : m.y) >> (16 * (k & 1));
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
            const uint32_t x = hlf ? (w[k] >> 16) : (w[k] & 0xffffu);
            const bool in = !MASKED || (mb & (hlf ? 0xff00u : 0xffu)) != 0u;
            if (PHASE == 0) {
                const uint32_t old = k2b_add(S, x, in);
                acc += S.gfix[old];                        // G[0] = 0: nothing for a pixel outside the mask
                maxold = max(maxold, old);
                cnt += in ? 1u : 0u;
            } else if (in) {
                S.hist[x >> 2] = 0u;                       // the whole word: a wrapped counter carried into its neighbour
            }
        }
    }
}

template <int PHASE, bool MASKED>
__device__ __forceinline__ void k2b_walk(K2bSmem& S, const Tile& T, const uint4* vreg, const uint2* mreg, uint32_t& cnt,
                                         uint32_t& maxold, unsigned long long& acc) {
    const int tid = threadIdx.x;
    const uint4* px4 = reinterpret_cast<const uint4*>(T.px);
    const uint2* mk2 = reinterpret_cast<const uint2*>(T.mk);
    const int nfull = T.n >> 3, rem = T.n & 7;
#pragma unroll
    for (int i = 0; i < kK2bVec; ++i)
        if (tid + i * kK2bThreads < nfull) k2b_vec<PHASE, MASKED>(S, vreg[i], mreg[i], cnt, maxold, acc);
    for (int idx = tid + kK2bVec * kK2bThreads; idx < nfull; idx += kK2bThreads) {
        const uint4 v = ld_reuse(px4 + idx);
        uint2 m = make_uint2(0u, 0u);
        if (MASKED) m = __ldg(mk2 + idx);
        k2b_vec<PHASE, MASKED>(S, v, m, cnt, maxold, acc);
    }
    if (tid < rem) {
        const int i = nfull * 8 + tid;
        const bool in = !MASKED || T.mk[i] != 0;
        const uint32_t x = T.px[i];
        if (PHASE == 0) {
            const uint32_t old = k2b_add(S, x, in);
            acc += S.gfix[old];
            maxold = max(maxold, old);
            cnt += in ? 1u : 0u;
        } else if (in) {
            S.hist[x >> 2] = 0u;
        }
    }
}

// numpy percentile (method "linear"): warp gw takes the percentiles k = gw, gw + 8 (see k2_percentiles)
__device__ __forceinline__ void k2b_percentiles(const K2bSmem& S, const Params& P, int n, double* o, int gw) {
    const int lane = threadIdx.x & 31;
    const uint4* c4 = reinterpret_cast<const uint4*>(S.c256);
    const uint4 qa = c4[2 * lane], qb = c4[2 * lane + 1];
    const int c[8] = {(int)qa.x, (int)qa.y, (int)qa.z, (int)qa.w, (int)qb.x, (int)qb.y, (int)qb.z, (int)qb.w};
    const int tot = c[0] + c[1] + c[2] + c[3] + c[4] + c[5] + c[6] + c[7];
    int incl = tot;
#pragma unroll
    for (int o2 = 1; o2 < 32; o2 <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o2);
        if (lane >= o2) incl += v;
    }
    const int excl = incl - tot;
    for (int k = gw; k < 9; k += kK2bWarps) {
        const double virt = __dmul_rn((double)(n - 1), P.quant[k]);
        int rk[2];
        if (virt >= (double)(n - 1)) { rk[0] = rk[1] = n - 1; }
        else { rk[0] = (int)floor(virt); rk[1] = rk[0] + 1; }
        int val[2], sl_prev = -1;
        int d[8] = {0, 0, 0, 0, 0, 0, 0, 0}, in2 = 0, tl = 0;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const uint32_t own = __ballot_sync(0xffffffffu, rk[j] >= excl && rk[j] < incl);
            const int src = __ffs(own) - 1;
            const int t1 = k2_locate8(c, rk[j] - excl);
            int below = excl;
#pragma unroll
            for (int u = 0; u < 7; ++u) below += u < t1 ? c[u] : 0;
            const int sl = __shfl_sync(0xffffffffu, 8 * lane + t1, src);
            const int base = __shfl_sync(0xffffffffu, below, src);
            if (sl != sl_prev) {                           // lane l holds the values 8l .. 8l+7 of the slice (8 byte counters)
                IMFEAT_CHECK(sl >= 0 && sl < 256 && own != 0u);
                const uint2 q = reinterpret_cast<const uint2*>(S.hist)[sl * 32 + lane];
                d[0] = (int)(q.x & 0xffu); d[1] = (int)((q.x >> 8) & 0xffu); d[2] = (int)((q.x >> 16) & 0xffu); d[3] = (int)(q.x >> 24);
                d[4] = (int)(q.y & 0xffu); d[5] = (int)((q.y >> 8) & 0xffu); d[6] = (int)((q.y >> 16) & 0xffu); d[7] = (int)(q.y >> 24);
                tl = d[0] + d[1] + d[2] + d[3] + d[4] + d[5] + d[6] + d[7];
                in2 = tl;
#pragma unroll
                for (int o2 = 1; o2 < 32; o2 <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, in2, o2);
                    if (lane >= o2) in2 += v;
                }
                sl_prev = sl;
            }
            const int r1 = base + in2, r0 = r1 - tl;
            const uint32_t own2 = __ballot_sync(0xffffffffu, rk[j] >= r0 && rk[j] < r1);
            IMFEAT_CHECK(own2 != 0u && (own2 & (own2 - 1u)) == 0u);
            const int t2 = k2_locate8(d, rk[j] - r0);
            val[j] = __shfl_sync(0xffffffffu, sl * 256 + 8 * lane + t2, __ffs(own2) - 1);
        }
        if (lane == 0) {
            const double g = virt - floor(virt);
            const int a = val[0], b = val[1];
            const double diff = (double)(b - a);
            // numpy _lerp: a + diff*t, replaced by b - diff*(1-t) where t >= 0.5 (no FMA there)
            o[1 + k] = (g >= 0.5) ? __dsub_rn((double)b, __dmul_rn(diff, __dsub_rn(1.0, g)))
                                  : __dadd_rn((double)a, __dmul_rn(diff, g));
        }
    }
}

// worklist / worklist_count: the tiles K12 left over; left / left_count: the tiles this kernel leaves to K2
// (a byte counter wrapped).  sched: the launch's tile counter.
template <bool MASKED>
__global__ void __launch_bounds__(kK2bThreads, 3)
k2b_order_entropy_kernel(const __grid_constant__ Params P, const uint32_t* __restrict__ worklist,
                         const uint32_t* __restrict__ worklist_count, uint32_t* __restrict__ left,
                         uint32_t* __restrict__ left_count, unsigned int* __restrict__ sched) {
    extern __shared__ __align__(16) unsigned char k2b_smem_raw[];
    K2bSmem& S = *reinterpret_cast<K2bSmem*>(k2b_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t total = *worklist_count;
    if (blockIdx.x >= total) return;
    for (int k = tid; k < 16384; k += kK2bThreads) S.hist[k] = 0u;
    S.c256[tid] = 0u;
    S.gfix[tid] = __ldg(P.gfix + tid);
    if (tid == 0) { S.cnt = 0u; S.maxold = 0u; S.next[0] = blockIdx.x; }
    __syncthreads();

    for (uint32_t it = 0;; ++it) {
        const uint32_t k = S.next[it & 1u];                // written two barriers ago at the latest
        if (k >= total) break;
        const uint32_t tile = worklist[k];
        const Tile T = resolve_tile(P, (long long)tile);
        double* o = T.out_row + P.col_basic + kNBasic * T.slot;
        uint4 vreg[kK2bVec];
        uint2 mreg[kK2bVec];
        {
            const uint4* px4 = reinterpret_cast<const uint4*>(T.px);
            const uint2* mk2 = reinterpret_cast<const uint2*>(T.mk);
            const int nfull = T.n >> 3;
#pragma unroll
            for (int i = 0; i < kK2bVec; ++i) {
                const int idx = tid + i * kK2bThreads;
                mreg[i] = make_uint2(0u, 0u);
                vreg[i] = make_uint4(0u, 0u, 0u, 0u);
                if (idx < nfull) {
                    vreg[i] = ld_stream(px4 + idx);
                    if (MASKED) mreg[i] = __ldg(mk2 + idx);
                }
            }
        }
        uint32_t cnt = 0u, maxold = 0u;
        unsigned long long acc = 0ull;
        k2b_walk<0, MASKED>(S, T, vreg, mreg, cnt, maxold, acc);
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        maxold = __reduce_max_sync(0xffffffffu, maxold);
        acc = warp_sum_redux(acc);
        if (lane == 0) {
            if (cnt) atomicAdd(&S.cnt, cnt);
            atomicMax(&S.maxold, maxold);
            S.wacc[warp] = acc;
        }
        if (tid == 0) S.next[(it + 1u) & 1u] = gridDim.x + atomicAdd(sched, 1u);   // next tile, read after the barriers below
        __syncthreads();                                   // ---- histogram complete ----
        const int n = (int)S.cnt;
        const uint32_t mo = S.maxold;
        const bool wrapped = mo >= 255u;                   // some bin reached 256: its counter wrapped
        if (!wrapped && n > 0) k2b_percentiles(S, P, n, o, warp);
        if (tid == 0) {
            if (wrapped) {
                left[atomicAdd(left_count, 1u)] = tile;    // K2's 16-bit table takes it
            } else if (n > 0) {
                unsigned long long tot = 0ull;
                for (int w = 0; w < kK2bWarps; ++w) tot += S.wacc[w];
                const double H = __ldg(P.log2tab + n) - ((double)tot * 2.2737367544323206e-13) / (double)n;
                o[16] = ((int)mo + 1 == n) ? 0.0 : H;      // one value only: entropy is exactly 0
            } else {
                const double nan = qnan();
#pragma unroll
                for (int q = 1; q <= 9; ++q) o[q] = nan;
                o[16] = nan;
            }
        }
        __syncthreads();                                   // ---- table read ----
        k2b_walk<2, MASKED>(S, T, vreg, mreg, cnt, maxold, acc);
        S.c256[tid] = 0u;
        if (tid == 0) { S.cnt = 0u; S.maxold = 0u; }
        __syncthreads();                                   // ---- table clean ----
    }
}

}  // namespace imfeat
