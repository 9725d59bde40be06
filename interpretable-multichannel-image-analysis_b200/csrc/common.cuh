// common.cuh -- shared device helpers for the feature kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Device-side bounds checks (-DIMFEAT_CHECKS, profiles/build_variant.sh): every shared-memory table, staging buffer
// and work-list index the kernels compute is tested and a violation traps the kernel -- the GPU test-suite run on
// such a build stands in for compute-sanitizer's memcheck, which is closed on this GPU pool (profiles/).
#ifdef IMFEAT_CHECKS
#include <stdio.h>
#define IMFEAT_CHECK(cond)                                                                   \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            printf("IMFEAT_CHECK failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__);           \
            __trap();                                                                        \
        }                                                                                    \
    } while (0)
#else
#define IMFEAT_CHECK(cond) ((void)0)
#endif

namespace imfeat {

constexpr int kNBasic = 17;
constexpr int kNGlcm = 6;
constexpr int kNShape = 10;
constexpr int kNMoment = 9;
constexpr int kMaxAngles = 4;
constexpr int kMaxPixels = 32768;
constexpr int kLevels = 256;

constexpr uint32_t kStEmptyMask = 1u;
constexpr uint32_t kStNoPairs = 2u;
constexpr uint32_t kStConstant = 4u;

// Everything a kernel needs to find its tiles and its output columns.  A "tile" is one
// (output row, channel slot) plane: tile t -> row t / c_out, slot t % c_out.
struct Params {
    const uint16_t* planes;
    const uint8_t* masks;      // nullable
    const int32_t* sizes;      // nullable, [N][2]
    const int32_t* src_obj;    // nullable, [N][c_out]
    const int32_t* chan;       // nullable, [c_out]
    double* out;
    uint32_t* status;          // nullable
    const double* log2tab;     // log2(k), k = 0..kMaxPixels (k=0 -> 0)
    unsigned int* sched;       // per-call work counters, one per kernel (dynamic tile scheduling)
    int k1_fp64_only;          // IMFEAT_K1_FP64=1: skip K1's integer pass (testing the fallback)
    const unsigned long long* gfix;  // round(2^42 * ((k+1)*log2(k+1) - k*log2(k))), k < kMaxPixels
    uint32_t* counts;          // nullable: GLCM bin dump [tile][angle][65536]
    long long n_tiles;
    long long plane_stride;
    long long row_stride;
    int c_in, c_out, hs, ws;
    int col_basic, col_glcm, col_shape, col_moment;  // block bases (col_* < 0: block absent)
    int n_angles;
    int dr[kMaxAngles], dc[kMaxAngles];
    double quant[9];           // percentile q / 100, computed on the host like numpy does
};

struct Tile {
    const uint16_t* px;
    const uint8_t* mk;
    double* out_row;
    uint32_t* status;
    int slot, h, w, n;
};

__device__ __forceinline__ Tile resolve_tile_rs(const Params& P, uint32_t row32, uint32_t slot32) {
    Tile T;
    const long long row = row32;
    const int slot = (int)slot32;
    const long long t = row * P.c_out + slot;
    const long long so = P.src_obj ? (long long)P.src_obj[t] : row;
    const int ch = P.chan ? P.chan[slot] : slot;
    T.h = P.sizes ? P.sizes[2 * so] : P.hs;
    T.w = P.sizes ? P.sizes[2 * so + 1] : P.ws;
    T.n = T.h * T.w;
    const long long off = (so * P.c_in + ch) * P.plane_stride;
    T.px = P.planes + off;
    T.mk = P.masks ? P.masks + off : nullptr;
    T.out_row = P.out + row * P.row_stride;
    T.status = P.status ? P.status + row : nullptr;
    T.slot = slot;
    return T;
}

__device__ __forceinline__ Tile resolve_tile(const Params& P, long long t) {
    // n_tiles < 2^31 is enforced by the host API: 32-bit division is much cheaper than 64-bit
    const uint32_t row32 = (uint32_t)t / (uint32_t)P.c_out;
    return resolve_tile_rs(P, row32, (uint32_t)t - row32 * (uint32_t)P.c_out);
}

// Walks tiles t0, t0 + stride, t0 + 2*stride, ... keeping (row, slot) = (t / c_out, t % c_out)
// incrementally, so the per-tile integer divisions disappear from the persistent loops.
struct TileWalk {
    uint32_t row, slot, drow, dslot, c_out;
    __device__ __forceinline__ void init(const Params& P, long long t0, long long stride) {
        c_out = (uint32_t)P.c_out;
        row = (uint32_t)t0 / c_out; slot = (uint32_t)t0 - row * c_out;
        drow = (uint32_t)stride / c_out; dslot = (uint32_t)stride - drow * c_out;
    }
    __device__ __forceinline__ void next() {
        row += drow; slot += dslot;
        if (slot >= c_out) { slot -= c_out; ++row; }
    }
    __device__ __forceinline__ long long tile() const { return (long long)row * c_out + slot; }
};

// Dynamic tile scheduler of the warp-per-tile kernels: every warp draws its next tile from a
// per-launch global counter (one atomic per tile, fetched one tile ahead so its latency is hidden).
// Unlike a static stride it needs no assumption about how many CTAs are resident at once -- e.g.
// while an NCCL all-gather of the previous batch occupies part of the machine.
__device__ __forceinline__ long long next_tile(unsigned int* counter) {
    unsigned int t = 0;
    if ((threadIdx.x & 31) == 0) t = atomicAdd(counter, 1u);
    return (long long)__shfl_sync(0xffffffffu, t, 0);
}

// A warp-per-tile kernel that draws its tiles two ahead can pull the next tile's pixels and mask bytes into L2 while
// the current tile is worked off.  Measured (round 2): no gain for K12 / K4w / K3a at 20-36 warps per SM (64x64 tiles:
// 2.98 vs 2.93 ms per step, the extra address arithmetic costs more than the shorter wait saves), but 9 % for the
// front kernel of K3 on 128x128 strides, where shared memory leaves it 8 warps per SM -- used there only.
__device__ __forceinline__ void prefetch_tile_l2(const Params& P, long long t, long long t_end) {
    if (t >= t_end) return;
    const Tile T = resolve_tile(P, t);
    const int lane = threadIdx.x & 31;
    const int lpx = (T.n * 2 + 127) >> 7, lmk = T.mk ? (T.n + 127) >> 7 : 0;      // 128-byte lines
    for (int l = lane; l < lpx + lmk; l += 32) {
        const char* line = l < lpx ? reinterpret_cast<const char*>(T.px) + 128 * l
                                   : reinterpret_cast<const char*>(T.mk) + 128 * (l - lpx);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
    }
}

// 128-bit streaming load: the tile is read once per kernel, keep it out of L1.
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
// 128-bit load that may be re-read by the same SM a few microseconds later (allocate in L1).
__device__ __forceinline__ uint4 ld_reuse(const uint4* p) { return __ldg(p); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Named barriers (ids 1..15; 0 is __syncthreads).  bar_arrive + bar_sync on the same id is the
// producer/consumer hand-off: the arriving threads' prior shared-memory writes are visible to the
// threads released by bar_sync.
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- Token ring over one shared-memory table (K2; K3 hands its table over with named barriers) -------------------------------------------
// The 65,536-bin table leaves room for one CTA per SM.  Its 1,024 threads are split into NG
// groups (2, 4 or 8) that work on NG different tiles; the table is handed round-robin from group
// to group through mbarriers: token[g] completes a phase when every thread of group g-1 has
// arrived (after its last table access), and group g waits on it before its first access.  All
// table-free work of a group overlaps the other groups' table phases.
struct Ring {
    int g, ng, gthreads, gt, gw, gwarps;
    uint32_t tok_mine, tok_next;   // shared-window addresses of token[g], token[(g+1) % ng]
    uint32_t parity;               // parity of the phase this group waits for next
};
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ring_init(Ring& R, unsigned long long* tokens, int ng) {
    const int tid = threadIdx.x;
    R.ng = ng;
    R.gthreads = blockDim.x / ng;
    R.g = tid / R.gthreads;
    R.gt = tid - R.g * R.gthreads;
    R.gw = R.gt >> 5;
    R.gwarps = R.gthreads >> 5;
    R.tok_mine = smem_addr(tokens + R.g);
    R.tok_next = smem_addr(tokens + (R.g + 1 == ng ? 0 : R.g + 1));
    R.parity = R.g == 0 ? 1u : 0u;     // group 0 owns the table first (phase "-1" counts as complete)
    if (tid < ng)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(tokens + tid)), "r"(R.gthreads));
}
__device__ __forceinline__ void ring_acquire(Ring& R) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_%=;\n"
        "}\n" ::"r"(R.tok_mine), "r"(R.parity)
        : "memory");
    R.parity ^= 1u;
}
__device__ __forceinline__ void ring_release(const Ring& R) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(R.tok_next) : "memory");
}
__device__ __forceinline__ void ring_group_sync(const Ring& R) { bar_sync(1 + R.g, R.gthreads); }

// ---- mbarrier + bulk async copy (the 1-D TMA path, UBLKCP) --------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@!p bra WAIT_%=;\n"
        "}\n" ::"r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// exact warp sum of 64-bit integers with three 21-bit limbs (REDUX.ADD is one instruction)
__device__ __forceinline__ unsigned long long warp_sum_redux(unsigned long long v) {
    const uint32_t l0 = (uint32_t)v & 0x1fffffu, l1 = (uint32_t)(v >> 21) & 0x1fffffu;
    const uint32_t l2 = (uint32_t)(v >> 42);
    const unsigned long long s0 = __reduce_add_sync(0xffffffffu, l0);
    const unsigned long long s1 = __reduce_add_sync(0xffffffffu, l1);
    const unsigned long long s2 = __reduce_add_sync(0xffffffffu, l2);
    return s0 + (s1 << 21) + (s2 << 42);
}

// The same three limb sums combined in floating point: the total of 32 lanes may exceed 64 bits even though
// every lane's value fits (K1's sum of fourth powers: 1,024 * y^4 < 2^64 per lane, times 32 lanes).
__device__ __forceinline__ double warp_sum_redux_dbl(unsigned long long v) {
    const uint32_t l0 = (uint32_t)v & 0x1fffffu, l1 = (uint32_t)(v >> 21) & 0x1fffffu;
    const uint32_t l2 = (uint32_t)(v >> 42);
    const uint32_t s0 = __reduce_add_sync(0xffffffffu, l0);         // 32 * 2^21 and 32 * 2^22 fit 32 bits
    const uint32_t s1 = __reduce_add_sync(0xffffffffu, l1);
    const uint32_t s2 = __reduce_add_sync(0xffffffffu, l2);
    return (double)s0 + (double)s1 * 2097152.0 + (double)s2 * 4398046511104.0;
}

__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7ff8000000000000LL); }

}  // namespace imfeat
