// k2_order_entropy.cuh -- K2: order statistics + Shannon entropy.
//
// Replaces, per channel:  np.percentile(X, 0.1 .. 0.9)  NB:242-250   shannon_entropy(X)  NB:262
//
// Both need the multiplicity of every raw 16-bit value of the tile.  The CTA owns one privatised
// 65,536-bin histogram in shared memory (16-bit counters, two per word, 128 KB); every thread group
// keeps, next to it, the 256 counts of its tile's values by high byte.  The groups work on different
// tiles and take turns on the table (common.cuh, token ring).
//   entropy      sum_values c*log2(c) telescopes over the atomics' return values:
//                sum_pixels G[old_p],  G[k] = (k+1)log2(k+1) - k*log2(k)   -> no read-back pass.
//                G is held in 2^-42 fixed point and summed in 64-bit integers, so the result
//                does not depend on the order in which the atomics resolve (bit-reproducible)
//   percentiles  two levels: one warp scans the 256 high-byte counts (one pass: which 256-value
//                slice holds which rank, and how many pixels lie below it), then reads only the slices
//                that hold a rank -- 1 KB of the table each, whatever the tile's value range -- and
//                applies numpy's linear-interpolation formula bit for bit.  (Round 1 walked the occupied
//                64-value blocks from the bottom: 30 dependent steps per tile on full-range data, 920 for
//                a median.)
//   clear        by re-walking the pixels (registers), never densely
#pragma once
#include "common.cuh"

namespace imfeat {

constexpr int kK2Vec = 2;      // 16-byte vectors per thread kept in registers (pixels + returned counts)

struct alignas(16) K2Group {
    uint32_t c256[256];      // pixels of the group's tile by value >> 8 (fire-and-forget atomics next to the table's)
    int vals[18];
    uint32_t cnt;            // masked pixel count (integer atomics)
    int constant;
    unsigned long long wacc[16];   // per-warp sums of G[old] in 2^-42 fixed point
};
constexpr int kK2GfixSmem = 4096;  // entropy-table entries mirrored in shared memory
struct K2Smem {
    uint32_t hist[32768];
    uint32_t dummy[32];      // one word per lane, right behind the table: where pixels outside the mask go
    unsigned long long gfix[kK2GfixSmem];
    unsigned long long tokens[8];
    K2Group grp[8];
};
__device__ __forceinline__ unsigned long long k2_gfix(const K2Smem& S, const Params& P, uint32_t old) {
    return old < (uint32_t)kK2GfixSmem ? S.gfix[old] : __ldg(P.gfix + old);
}

// returns the count of x BEFORE this increment
__device__ __forceinline__ uint32_t k2_add(K2Smem& S, uint32_t x) {
    const uint32_t sh = (x & 1u) << 4;
    const uint32_t old = atomicAdd(&S.hist[x >> 1], 1u << sh);
    return (old >> sh) & 0xffffu;
}

// PHASE 0: histogram build + entropy terms, 2: sparse clear.
// olds (optional): the returned counts of the 8 pixels, packed 2 x 16 bit per word; when given,
// the entropy-table look-ups are left to the caller (after the table has been handed over).
template <int PHASE, bool MASKED>
__device__ __forceinline__ void k2_vec(K2Smem& S, uint32_t* c256, const Params& P, const uint4& v, const uint2& m,
                                       uint32_t& cnt, uint32_t& maxold, unsigned long long& acc,
                                       uint32_t* olds = nullptr) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    if (PHASE == 0 && olds) { olds[0] = olds[1] = olds[2] = olds[3] = 0u; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t mb = (k < 2 ? m.x : m.y) >> (16 * (k & 1));
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
            const uint32_t x = hlf ? (w[k] >> 16) : (w[k] & 0xffffu);
            if (!MASKED) {
                if (PHASE == 0) {
                    const uint32_t old = k2_add(S, x);
                    atomicAdd(c256 + (x >> 8), 1u);
                    if (olds) {
                        olds[k] |= old << (16 * hlf);
                    } else {
                        acc += k2_gfix(S, P, old);
                        maxold = max(maxold, old);
                    }
                } else {
                    S.hist[x >> 1] = 0u;
                }
            } else {
                // branch-free: a pixel outside the mask increments (and later clears) this lane's
                // dummy word instead of a bin, and its returned count is replaced by 0 (G[0] = 0)
                const bool in = (mb & (hlf ? 0xff00u : 0xffu)) != 0u;
                const uint32_t off = in ? ((x << 1) & 0x1fffcu) : (0x20000u + 4u * (threadIdx.x & 31));
                uint32_t* word = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(S.hist) + off);
                if (PHASE == 0) {
                    const uint32_t sh = in ? ((x & 1u) << 4) : 0u;
                    uint32_t old = (atomicAdd(word, 1u << sh) >> sh) & 0xffffu;
                    atomicAdd(c256 + (x >> 8), in ? 1u : 0u);
                    old = in ? old : 0u;
                    cnt += in ? 1u : 0u;
                    if (olds) {
                        olds[k] |= old << (16 * hlf);
                    } else {
                        acc += k2_gfix(S, P, old);
                        maxold = max(maxold, old);
                    }
                } else {
                    *word = 0u;
                }
            }
        }
    }
}

template <int PHASE, bool MASKED>
__device__ __forceinline__ void k2_walk(K2Smem& S, uint32_t* c256, const Params& P, const Tile& T, int gt, int gthreads,
                                        const uint4* vreg, const uint2* mreg, uint32_t& cnt,
                                        uint32_t& maxold, unsigned long long& acc, uint32_t (*olds)[4] = nullptr) {
    const uint4* px4 = reinterpret_cast<const uint4*>(T.px);
    const uint2* mk2 = reinterpret_cast<const uint2*>(T.mk);
    const int nfull = T.n >> 3, rem = T.n & 7;
#pragma unroll
    for (int i = 0; i < kK2Vec; ++i)
        if (gt + i * gthreads < nfull)
            k2_vec<PHASE, MASKED>(S, c256, P, vreg[i], mreg[i], cnt, maxold, acc, (PHASE == 0 && olds) ? olds[i] : nullptr);
    for (int idx = gt + kK2Vec * gthreads; idx < nfull; idx += gthreads) {
        const uint4 v = ld_reuse(px4 + idx);
        uint2 m = make_uint2(0u, 0u);
        if (MASKED) m = __ldg(mk2 + idx);
        k2_vec<PHASE, MASKED>(S, c256, P, v, m, cnt, maxold, acc);
    }
    if (gt < rem) {
        const int i = nfull * 8 + gt;
        if (!MASKED || T.mk[i] != 0) {
            const uint32_t x = T.px[i];
            if (PHASE == 0) {
                const uint32_t old = k2_add(S, x);
                atomicAdd(c256 + (x >> 8), 1u);
                acc += k2_gfix(S, P, old);
                maxold = max(maxold, old);
                if (MASKED) ++cnt;
            } else {
                S.hist[x >> 1] = 0u;
            }
        }
    }
}

// numpy percentile (method "linear") from the histogram; ranks are 0-based positions in the sorted
// multiset.  Every warp of the group takes the percentiles k = gw, gw + gwarps, ..: level 1 on the
// high-byte counts (each warp for itself: two 128-bit loads and one scan), then one 1 KB slice of the table
// per rank.  No shared scratch: the lane that holds a rank hands its answer round with shuffles.
__device__ __forceinline__ int k2_locate8(const int (&c)[8], int rel) {      // t with sum(c[<t]) <= rel < sum(c[<=t])
    int t = 0, run = 0;
#pragma unroll
    for (int u = 0; u < 7; ++u) { run += c[u]; t += rel >= run ? 1 : 0; }
    return t;
}
__device__ __forceinline__ void k2_percentiles(const K2Smem& S, const K2Group& G, const Params& P, int n, double* o,
                                               int gw, int gwarps) {
    const int lane = threadIdx.x & 31;
    // level 1: lane l holds the counts of the slices 8l .. 8l+7
    const uint4* c4 = reinterpret_cast<const uint4*>(G.c256);
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
    for (int k = gw; k < 9; k += gwarps) {
        const double virt = __dmul_rn((double)(n - 1), P.quant[k]);
        int rk[2];
        if (virt >= (double)(n - 1)) { rk[0] = rk[1] = n - 1; }
        else { rk[0] = (int)floor(virt); rk[1] = rk[0] + 1; }
        int val[2], sl_prev = -1;
        int d[8] = {0, 0, 0, 0, 0, 0, 0, 0}, in2 = 0, tl = 0;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            // the lane whose eight slices hold the rank
            const uint32_t own = __ballot_sync(0xffffffffu, rk[j] >= excl && rk[j] < incl);
            const int src = __ffs(own) - 1;
            const int t1 = k2_locate8(c, rk[j] - excl);
            int below = excl;
#pragma unroll
            for (int u = 0; u < 7; ++u) below += u < t1 ? c[u] : 0;
            const int sl = __shfl_sync(0xffffffffu, 8 * lane + t1, src);
            const int base = __shfl_sync(0xffffffffu, below, src);       // pixels in the slices below
            if (sl != sl_prev) {                           // level 2: lane l holds the values 8l .. 8l+7 of the slice
                IMFEAT_CHECK(sl >= 0 && sl < 256 && own != 0u);
                const uint4 q = reinterpret_cast<const uint4*>(S.hist)[sl * 32 + lane];
                d[0] = (int)(q.x & 0xffffu); d[1] = (int)(q.x >> 16); d[2] = (int)(q.y & 0xffffu); d[3] = (int)(q.y >> 16);
                d[4] = (int)(q.z & 0xffffu); d[5] = (int)(q.z >> 16); d[6] = (int)(q.w & 0xffffu); d[7] = (int)(q.w >> 16);
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
            IMFEAT_CHECK(own2 != 0u && (own2 & (own2 - 1u)) == 0u);          // exactly one lane holds the rank
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

template <bool MASKED>
__global__ void __launch_bounds__(1024, 1) k2_order_entropy_kernel(const __grid_constant__ Params P, int ng,
                                                                   const uint32_t* __restrict__ worklist,
                                                                   const uint32_t* __restrict__ worklist_count) {
    extern __shared__ __align__(16) unsigned char k2_smem_raw[];
    K2Smem& S = *reinterpret_cast<K2Smem*>(k2_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    // nothing on the worklist for this CTA (12-bit data: nothing at all): leave before the 160 KB of tables are set up
    if (worklist && (long long)blockIdx.x >= (long long)*worklist_count) return;
    Ring R;
    ring_init(R, S.tokens, ng);
    const int g = R.g, gt = R.gt, gw = R.gw, gthreads = R.gthreads;
    K2Group& G = S.grp[g];

    for (int k = tid; k < 32768; k += blockDim.x) S.hist[k] = 0u;
    for (int k = gt; k < 256; k += gthreads) G.c256[k] = 0u;
    if (tid < 32) S.dummy[tid] = 0u;
    if (gt == 0) { G.constant = 0; G.cnt = 0u; }
    for (int k = tid; k < kK2GfixSmem; k += blockDim.x) S.gfix[k] = __ldg(P.gfix + k);
    __syncthreads();

    // With a worklist (tiles the compact kernel k2c could not take: value range >= 4096) only those
    // tiles are processed; otherwise all tiles.
    const long long total = worklist ? (long long)*worklist_count : P.n_tiles;
    const long long first = blockIdx.x;
    const long long mine = first < total ? (total - first + gridDim.x - 1) / gridDim.x : 0;
    const long long n_iter = (mine + ng - 1) / ng;
    TileWalk walk;
    walk.init(P, first + (long long)g * gridDim.x < total ? first + (long long)g * gridDim.x : 0,
              (long long)ng * gridDim.x);
    auto resolve_k = [&](long long k) -> Tile {      // k-th tile of this CTA (k = ng*it + g)
        if (worklist) return resolve_tile(P, (long long)worklist[first + k * gridDim.x]);
        return resolve_tile_rs(P, walk.row, walk.slot);
    };

    // The pixels of a tile live in registers; the next tile's loads are issued right after the
    // table has been handed over, so HBM latency hides behind the current tile's epilogue and the
    // other groups' table phases.
    uint4 vreg[kK2Vec];
    uint2 mreg[kK2Vec];
    Tile T;
    bool active = g < mine;
    auto fetch = [&](const Tile& Tn) {
        const uint4* px4 = reinterpret_cast<const uint4*>(Tn.px);
        const uint2* mk2 = reinterpret_cast<const uint2*>(Tn.mk);
        const int nfull = Tn.n >> 3;
#pragma unroll
        for (int i = 0; i < kK2Vec; ++i) {
            const int idx = gt + i * gthreads;
            mreg[i] = make_uint2(0u, 0u);
            if (idx < nfull) {
                vreg[i] = ld_stream(px4 + idx);
                if (MASKED) mreg[i] = __ldg(mk2 + idx);
            }
        }
    };
    if (active) { T = resolve_k(g); fetch(T); }

    for (long long it = 0; it < n_iter; ++it) {
        uint32_t cnt = 0, maxold = 0;
        uint32_t olds[kK2Vec][4];
        unsigned long long acc = 0ull;
        double* o = nullptr;
        if (active) {
            o = T.out_row + P.col_basic + kNBasic * T.slot;
            // the tile's pixels must have landed BEFORE the table is taken: a group that waits for HBM while it
            // holds the table stalls the three others (an empty asm that reads a register waits for its load)
#pragma unroll
            for (int i = 0; i < kK2Vec; ++i) {
                asm volatile("" ::"r"(vreg[i].x), "r"(vreg[i].y), "r"(vreg[i].z), "r"(vreg[i].w));
                if (MASKED) asm volatile("" ::"r"(mreg[i].x), "r"(mreg[i].y));
            }
        }
        ring_acquire(R);                                   // ---- table owned by this group ----
        int n = 0;
        if (active) {
            k2_walk<0, MASKED>(S, G.c256, P, T, gt, gthreads, vreg, mreg, cnt, maxold, acc, olds);
            if (MASKED) {
                cnt = __reduce_add_sync(0xffffffffu, cnt);
                if (lane == 0 && cnt) atomicAdd(&G.cnt, cnt);
            }
            ring_group_sync(R);                            // histogram complete
            n = MASKED ? (int)G.cnt : T.n;
            if (n > 0) k2_percentiles(S, G, P, n, o, gw, R.gwarps);
            ring_group_sync(R);                            // percentiles done, n read by everyone
            k2_walk<2, MASKED>(S, G.c256, P, T, gt, gthreads, vreg, mreg, cnt, maxold, acc);
            for (int k = gt; k < 256; k += gthreads) G.c256[k] = 0u;
            if (gt == 0) G.cnt = 0u;
        }
        ring_release(R);                                   // ---- hand the table to the next group ----

        // prefetch the next tile of this group (the pixel registers are free again)
        const Tile Tcur = T;
        const bool was_active = active;
        walk.next();
        active = (long long)ng * (it + 1) + g < mine;
        if (active) { T = resolve_k((long long)ng * (it + 1) + g); fetch(T); }
        if (worklist && (long long)ng * (it + 2) + g < mine) {
            // the tile after next: pull it into L2 now (worklist tiles come in no order; their first touch is HBM latency)
            const Tile T2 = resolve_tile(P, (long long)worklist[first + ((long long)ng * (it + 2) + g) * gridDim.x]);
            const int lpx = (T2.n * 2 + 127) >> 7, lmk = MASKED ? (T2.n + 127) >> 7 : 0;     // 128-byte lines
            for (int l = gt; l < lpx + lmk; l += gthreads) {
                const char* line = l < lpx ? reinterpret_cast<const char*>(T2.px) + 128 * l
                                           : reinterpret_cast<const char*>(T2.mk) + 128 * (l - lpx);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
            }
        }

        if (was_active) {
            // entropy terms of the register-resident pixels: G[old], looked up off the critical path
            const int nfull = Tcur.n >> 3;
#pragma unroll
            for (int i = 0; i < kK2Vec; ++i)
                if (gt + i * gthreads < nfull) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t o2 = olds[i][q];
                        acc += k2_gfix(S, P, o2 & 0xffffu);
                        acc += k2_gfix(S, P, o2 >> 16);
                        maxold = max(maxold, max(o2 & 0xffffu, o2 >> 16));
                    }
                }
            if (n > 0 && (int)maxold + 1 == n) G.constant = 1;   // one value only: entropy is exactly 0
            acc = warp_sum_redux(acc);
            if (lane == 0) G.wacc[gw] = acc;
            ring_group_sync(R);
            if (gt == 0) {
                if (n > 0) {
                    // sum_values c*log2(c) = acc * 2^-42 ; H = log2 n - that / n
                    unsigned long long tot = 0ull;
                    for (int w = 0; w < R.gwarps; ++w) tot += G.wacc[w];
                    const double H = __ldg(P.log2tab + n) - ((double)tot * 2.2737367544323206e-13) / (double)n;
                    o[16] = G.constant ? 0.0 : H;
                } else {
                    const double nan = qnan();
#pragma unroll
                    for (int q = 1; q <= 9; ++q) o[q] = nan;
                    o[16] = nan;
                }
                G.constant = 0;
            }
        }
    }
}

// -------------------------------------------------------------------------------------------------
// K2c: compact variant for tiles whose value range (max - min, both already in the table from K1)
// is below 4,096 -- all 12-bit data.  The histogram is then 4,096 bins relative to the minimum
// (8 KB), private to a 256-thread CTA, so several CTAs run per SM with no table hand-over at all;
// the percentile walk starts at bin 0 (= the minimum) and the clear is a dense store over the used
// range.  Tiles with a wider range are appended to a worklist for the full-range ring kernel above.
// -------------------------------------------------------------------------------------------------
constexpr int kK2cThreads = 32;
constexpr int kK2cWarps = kK2cThreads / 32;
constexpr int kK2cBins = 4096;
constexpr int kK2cWords = kK2cBins / 2;

struct K2cSmem {
    uint32_t hist[kK2cWords];
    uint32_t dummy[32];
    int vals[18];
    uint32_t cnt;
    int constant;
    unsigned long long wacc[kK2cWarps];
};

// One pixel into the warp's histogram (fire-and-forget: the counts are read back once, by the pass that
// also clears them).  inb = all ones when the pixel counts (inside the mask), else 0: then the increment
// is 0, on word `lane` (distinct banks; equal background values would otherwise serialise on one address),
// so the masked path needs no branch and no dummy word.
template <bool MASKED>
__device__ __forceinline__ void k2c_px(K2cSmem& S, uint32_t x, uint32_t inb, uint32_t vmin) {
    const uint32_t bin = x - vmin;
    uint32_t off = (bin << 1) & (uint32_t)(kK2cWords * 4 - 4);
    uint32_t inc = (bin & 1u) * 0xffffu + 1u;
    if (MASKED) { inc &= inb; off = (off & inb) | (4u * (threadIdx.x & 31) & ~inb); }
    atomicAdd(reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(S.hist) + off), inc);
}

template <bool MASKED>
__global__ void __launch_bounds__(kK2cThreads, 24) k2c_order_entropy_kernel(const __grid_constant__ Params P,
                                                                          uint32_t* __restrict__ worklist,
                                                                          uint32_t* __restrict__ worklist_count) {
    __shared__ K2cSmem S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < kK2cWords; k += kK2cThreads) S.hist[k] = 0u;
    if (tid < 32) S.dummy[tid] = 0u;
    if (tid == 0) { S.cnt = 0u; S.constant = 0; }
    __syncthreads();

    static_assert(kK2cThreads == 32, "the dynamic scheduler below assumes one warp per CTA");
    // min and max of a tile (written by K1, earlier launch on the same stream), fetched one tile ahead
    auto peek = [&](long long tt, double& a, double& b) {
        if (tt >= P.n_tiles) return;
        const uint32_t row = (uint32_t)tt / (uint32_t)P.c_out, slot = (uint32_t)tt - row * (uint32_t)P.c_out;
        const double* q = P.out + (long long)row * P.row_stride + P.col_basic + kNBasic * (int)slot;
        a = q[0]; b = q[10];
    };
    long long tnext = next_tile(P.sched + 1);
    double dmin_n = 0.0, dmax_n = 0.0;
    peek(tnext, dmin_n, dmax_n);
    while (tnext < P.n_tiles) {
        const long long t = tnext;
        const double dmin = dmin_n, dmax = dmax_n;
        tnext = next_tile(P.sched + 1);                      // one tile ahead
        peek(tnext, dmin_n, dmax_n);
        const Tile T = resolve_tile(P, t);
        double* o = T.out_row + P.col_basic + kNBasic * T.slot;
        if (!(dmin == dmin)) {                               // NaN: no pixel inside the mask
            if (tid == 0) {
                const double nan = qnan();
#pragma unroll
                for (int q = 1; q <= 9; ++q) o[q] = nan;
                o[16] = nan;
            }
            continue;
        }
        const uint32_t vmin = (uint32_t)dmin, range = (uint32_t)dmax - vmin;
        if (range >= (uint32_t)kK2cBins) {                   // too wide: leave it to the full-range kernel
            if (tid == 0) worklist[atomicAdd(worklist_count, 1u)] = (uint32_t)t;
            continue;
        }
        const uint4* px4 = reinterpret_cast<const uint4*>(T.px);
        const uint2* mk2 = reinterpret_cast<const uint2*>(T.mk);
        const int nfull = T.n >> 3, rem = T.n & 7;
        uint32_t cnt = 0;
        // software pipeline: the loads of the next two chunks are in flight while this one goes into the histogram
        uint4 vn = make_uint4(0u, 0u, 0u, 0u), vn2 = vn;
        uint2 mn = make_uint2(0u, 0u), mn2 = mn;
        if (tid < nfull) { vn = ld_stream(px4 + tid); if (MASKED) mn = __ldg(mk2 + tid); }
        if (tid + kK2cThreads < nfull) { vn2 = ld_stream(px4 + tid + kK2cThreads); if (MASKED) mn2 = __ldg(mk2 + tid + kK2cThreads); }
        for (int idx = tid; idx < nfull; idx += kK2cThreads) {
            const uint4 v = vn;
            const uint2 m = mn;
            vn = vn2; mn = mn2;
            if (idx + 2 * kK2cThreads < nfull) {
                vn2 = ld_stream(px4 + idx + 2 * kK2cThreads);
                if (MASKED) mn2 = __ldg(mk2 + idx + 2 * kK2cThreads);
            }
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            const uint32_t nz[2] = {__vcmpne4(m.x, 0u), __vcmpne4(m.y, 0u)};      // 0xff per pixel inside the mask
            if (MASKED) cnt += (__popc(nz[0]) + __popc(nz[1])) >> 3;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t n4 = nz[k >> 1];
                k2c_px<MASKED>(S, w[k] & 0xffffu, __byte_perm(n4, 0u, (k & 1) ? 0x2222 : 0x0000), vmin);
                k2c_px<MASKED>(S, w[k] >> 16, __byte_perm(n4, 0u, (k & 1) ? 0x3333 : 0x1111), vmin);
            }
        }
        if (tid < rem) {
            const int i = nfull * 8 + tid;
            const bool in = !MASKED || T.mk[i] != 0;
            if (in) {
                k2c_px<false>(S, T.px[i], 0xffffffffu, vmin);
                cnt += 1;
            }
        }
        if (MASKED) {
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (lane == 0 && cnt) atomicAdd(&S.cnt, cnt);
        }
        __syncthreads();                                     // histogram complete
        const int n = MASKED ? (int)S.cnt : T.n;
        if (warp == 0 && n > 0) {
            // numpy percentile (method "linear"); dense walk from bin 0 = the minimum
            int lo[9], hi[9], maxrank = 0;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const double virt = __dmul_rn((double)(n - 1), P.quant[k]);
                if (virt >= (double)(n - 1)) { lo[k] = hi[k] = n - 1; }
                else { lo[k] = (int)floor(virt); hi[k] = lo[k] + 1; }
                maxrank = max(maxrank, hi[k]);
            }
            int cum = 0;
            for (int block = 0; block * 64 <= (int)range; ++block) {
                const uint32_t wv = S.hist[block * 32 + lane];
                const int c0 = wv & 0xffffu, c1 = wv >> 16, tot = c0 + c1;
                int incl = tot;
#pragma unroll
                for (int o2 = 1; o2 < 32; o2 <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o2);
                    if (lane >= o2) incl += v;
                }
                const int r0 = cum + incl - tot, r1 = cum + incl;
                const int base = (int)vmin + block * 64 + 2 * lane;
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    if (lo[k] >= r0 && lo[k] < r1) S.vals[2 * k] = base + (lo[k] >= r0 + c0);
                    if (hi[k] >= r0 && hi[k] < r1) S.vals[2 * k + 1] = base + (hi[k] >= r0 + c0);
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
                o[1 + lane] = (g >= 0.5) ? __dsub_rn((double)b, __dmul_rn(diff, __dsub_rn(1.0, g)))
                                         : __dadd_rn((double)a, __dmul_rn(diff, g));
            }
        }
        __syncthreads();                                     // walk done; n read
        // entropy = log2 n - (1/n) sum_bins c log2 c, in the pass that clears the used range: fixed order
        // per lane and a fixed shuffle tree, so the double sum is reproducible bit for bit
        const int used_words = (int)(range >> 1) + 1;
        double hs = 0.0;
        for (int k = tid; k < used_words; k += kK2cThreads) {
            const uint32_t wv = S.hist[k];
            if (wv) {
                const uint32_t c0 = wv & 0xffffu, c1 = wv >> 16;
                hs = fma((double)c0, __ldg(P.log2tab + c0), hs);       // log2tab[0] = 0
                hs = fma((double)c1, __ldg(P.log2tab + c1), hs);
                S.hist[k] = 0u;
            }
        }
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) hs += __shfl_xor_sync(0xffffffffu, hs, o2);
        if (tid == 0) {
            if (n > 0) {
                const double H = __ldg(P.log2tab + n) - hs / (double)n;
                o[16] = range == 0u ? 0.0 : H;               // one value only: entropy is exactly 0
            } else {
                const double nan = qnan();
#pragma unroll
                for (int q = 1; q <= 9; ++q) o[q] = nan;
                o[16] = nan;
            }
            S.cnt = 0u;
        }
        __syncthreads();                                     // table clean, scratch consumed
    }
}

}  // namespace imfeat
