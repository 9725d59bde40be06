// k3_ring.cuh -- K3 for UNMASKED tiles: gray-level quantisation + GLCM + Haralick properties in one kernel,
// one 128 KB table of 16-bit counters per SM handed round four thread groups ("ring").  Masked tiles go
// through the front / bins pair of kernels in k3_glcm.cuh, which is faster there (measured: 2.10 vs 2.54 ms
// per 10,000 masked objects x 4 directions); for unmasked tiles this kernel is (1.02 vs 1.30 ms): while one
// group works the table (shared-memory pipe) the other three do pair sums (ALU) on the same SM.
//
// Replaces, per channel:
//   (x / x.max()) * 255 -> uint8                      NB:293-295
//   greycomatrix(q, [5], [0], levels=256)              NB:298   (not symmetric, not normed)
//   greycoprops x6 (contrast .. correlation)           NB:301-306
//
// * quantiser: floor(255*x/max) with an exact multiply-shift reciprocal (bit-identical to the
//   notebook's float64 expression for every uint16 pair; tests/test_oracle_cpu.py).
// * the 256x256 bins live in shared memory as 16-bit counters (two per 32-bit word, 128 KB), built
//   with shared-memory atomics on the pair stream, dumped on request (parity), and only ever
//   cleared sparsely by re-walking the pairs.  One persistent CTA per SM; its four 256-thread groups
//   work on four tiles at a time and take turns on the table, handing it over with named barriers
//   (bar.arrive / bar.sync: the waiting group is parked in hardware and issues nothing).  While one
//   group owns the table the others load the pair items of their next direction, add up the pair-stream
//   sums (contrast, dissimilarity, homogeneity, correlation need no bins) and turn the pairs into hits.
// * each group stages its own tile (maximum from K1's column when the basic block ran, mask bits and
//   their bounding box, 8-bit quantisation into shared memory) while other groups use the table.  The
//   raw pixels and mask bytes of a group's next tile are prefetched into shared memory with cp.async
//   while it works on the current one, so staging never waits for global memory.  Per-tile scalar work
//   sits on one lane of the last warp; the epilogues of 8 parked tiles run at once on that warp.
// * ASM = sum_bins c^2 is accumulated from the atomics' return values
//   (c^2 = sum_{k<c} (2k+1) = 2*sum(old) + c), so there is no pass over the bins.
#pragma once
#include "common.cuh"

namespace imfeat {
namespace ring {

constexpr int kK3Threads = 1024;     // NG groups of 1024 / NG threads
constexpr int kK3Park = 8;           // finished tiles per group whose epilogues run side by side (x directions <= 32 lanes)

// per-group staging buffers behind K3Smem: quantised pixels (one byte each, row-major, + slack for the
// 16-byte item reads) and mask bits (masked variant only)
__host__ __device__ inline int k3_q8_words(int max_pixels) { return (max_pixels / 4 + 8 + 3) & ~3; }
__host__ __device__ inline int k3_mb_words(int max_pixels, bool masked) { return masked ? ((max_pixels / 32 + 2 + 3) & ~3) : 0; }
__host__ __device__ inline size_t k3_group_bytes(int max_pixels, bool masked) {
    return 4 * (size_t)(k3_q8_words(max_pixels) + k3_mb_words(max_pixels, masked));
}
// prefetch buffer of a group: the next tile's raw pixels, then its mask bytes (max_pixels is a multiple of 8)
__host__ __device__ inline size_t k3_raw_bytes(int max_pixels, bool masked) {
    return ((size_t)max_pixels * 2 + (masked ? (size_t)max_pixels : 0) + 15) & ~(size_t)15;
}

// What a group notes about its next tile while the current one is under way.
struct K3TileInfo {
    int h, w;
    uint32_t row, slot;              // output row, channel slot
    uint32_t mul, sh, fast;          // quantiser constants (k3_magic / k3_magic_fast) when K1's maximum is known
    int n16, n8;                     // 16-byte pixel chunks / 8-byte mask chunks to prefetch
    int pad;
    const uint16_t* px;              // where the prefetch copies read
    const uint8_t* mk;
};
// per group and direction, summed over the warps of the group
struct K3AccS {
    uint32_t s[8];                   // si sj sii sjj sij sd sold m
    uint32_t hom_lo, hom_hi;         // sum of 1 / (1 + d^2) in 2^-40 fixed point (64-bit shared atomics are CAS loops)
    uint32_t np, pad;                // pairs walked (16 per item), existing or not
};
struct alignas(16) K3Smem {                    // the staging buffers behind it hold 16-byte vectors
    uint32_t hist[32768];
    double homtab[256];                         // 1 / (1 + d^2)
    K3AccS acc[4][2][kK3Park][kMaxAngles];      // per group, bank, parking slot and direction: finished tiles wait
                                                // here until kK3Park of them get their epilogues at once, one per
                                                // lane, while the next batch already fills the other bank
    int box[4][2][4];                           // mask bounding box per group and tile parity: rmin rmax cmin cmax
    uint32_t wmax[4][32];                       // per-warp maxima (only when K1 did not run)
    K3TileInfo info[4][2];                      // per group and tile parity
    uint32_t tileid[4][2][kK3Park][2];          // output row and channel slot of the parked tiles
};
struct K3Group {                               // where the quantised tile and its mask bits live
    uint32_t* q8;                              // quantised pixels (bytes) + slack for unaligned reads
    uint32_t* mbits;                           // one bit per pixel: inside the mask (masked variant)
};
// groups sharing the table: 4, or 2 when four staging buffers do not fit
__host__ __device__ inline int k3_groups(int max_pixels, bool masked) {
    return sizeof(K3Smem) + 4 * k3_group_bytes(max_pixels, masked) <= 227 * 1024 ? 4 : 2;
}
// whether the raw-tile prefetch buffers fit next to the table
__host__ __device__ inline bool k3_prefetch(int max_pixels, bool masked) {
    const int ng = k3_groups(max_pixels, masked);
    return sizeof(K3Smem) + ng * (k3_group_bytes(max_pixels, masked) + k3_raw_bytes(max_pixels, masked)) <= 227 * 1024;
}
__host__ __device__ inline size_t k3_smem_bytes(int max_pixels, bool masked) {
    const int ng = k3_groups(max_pixels, masked);
    return sizeof(K3Smem) + (size_t)ng * (k3_group_bytes(max_pixels, masked) +
                                          (k3_prefetch(max_pixels, masked) ? k3_raw_bytes(max_pixels, masked) : 0));
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct K3Acc {
    uint32_t si, sj, sii, sjj, sij, sd, m;
    double hom;
};

// exact floor(255*x / vmax) for 0 <= x <= vmax <= 65535:  (255*x * mul) >> sh,
// mul = ceil(2^sh / vmax), sh = 24 + ceil(log2 vmax)
__device__ __forceinline__ void k3_magic(uint32_t vmax, uint32_t& mul, uint32_t& sh) {
    if (vmax == 0) { mul = 0; sh = 24; return; }
    const uint32_t l = (vmax <= 1) ? 0u : 32u - (uint32_t)__clz(vmax - 1);
    sh = 24u + l;
    const unsigned long long two = 1ull << sh;
    // float estimate (24-bit) then exact integer correction
    uint32_t m = (uint32_t)(__uint2float_rz(1u << l) * 16777216.0f / __uint2float_rn(vmax));
    while ((unsigned long long)m * vmax < two) ++m;
    while ((unsigned long long)(m - 1) * vmax >= two) --m;
    mul = m;
}
// 256 < vmax <= 4103: floor(255*x / vmax) = umulhi(x, 255 * ceil(2^32 / vmax)) for 0 <= x <= vmax
// (255 * vmax^2 < 2^32 bounds the rounding error; checked exhaustively in tests/test_oracle_cpu.py)
__device__ __forceinline__ bool k3_magic_fast(uint32_t vmax, uint32_t& mul) {
    if (vmax <= 256u || vmax > 4103u) return false;
    mul = 255u * (0xffffffffu / vmax + 1u);       // vmax is not a power of two here or the +1 is harmless: see test
    return true;
}
__device__ __forceinline__ uint32_t k3_quant(uint32_t x, uint32_t mul, uint32_t sh) {
    return (uint32_t)(((unsigned long long)(x * 255u) * mul) >> sh);
}

// four mask bits expanded to 0xff / 0x00 bytes
__device__ __forceinline__ uint32_t k3_expand4(uint32_t bits) {
    return (((bits & 0xfu) * 0x00204081u) & 0x01010101u) * 0xffu;
}
// 16 consecutive mask bits starting at bit offset off
__device__ __forceinline__ uint32_t k3_bits16(const uint32_t* b, int off) {
    const int w = off >> 5;
    return __funnelshift_r(b[w], b[w + 1], off & 31) & 0xffffu;
}

// An item is a run of up to 16 horizontally consecutive pairs (4 groups of 4) of one row: the words
// of both pixel runs are loaded once and funnel-shifted into place.  When the row length is a
// multiple of 16 the runs of one side (I for dc >= 0, J for dc < 0) are made to start at multiples
// of 16 bytes: that side is one conflict-free 128-bit load, the other side two.
struct K3Geom {
    int nrows, r0, c0, c1, ipr, items, w, doff;
    float rcp;                                     // 1 / ipr: row = floor((item + 0.5) * rcp), exact for item < 2^20
    int aligned, base_j, ws, sb;                   // aligned path: which side is 16-byte aligned; word / bit shift of the other
};
// Pairs (r, c) -> (r + dr, c + dc) with both pixels inside the box rows [br0, br1], columns
// [bc0, bc1] (the whole tile, or the bounding box of the mask: pairs outside it cannot exist).
template <bool MASKED>
__device__ __forceinline__ K3Geom k3_geom(int w, int dr, int dc, int br0, int br1, int bc0, int bc1) {
    K3Geom G;
    G.r0 = br0;
    G.nrows = br1 - dr - br0 + 1;                  // dr >= 0 for all supported directions
    G.c0 = bc0 + (dc < 0 ? -dc : 0);
    G.c1 = bc1 + 1 - (dc > 0 ? dc : 0);
    G.w = w;
    G.doff = dr * w + dc;
    G.aligned = 0; G.base_j = dc < 0; G.ws = 0; G.sb = 0;
    if (G.nrows <= 0 || G.c1 <= G.c0) { G.items = 0; G.ipr = 1; G.rcp = 1.0f; G.nrows = 0; return G; }
    // unmasked tiles only: the aligned side starts at column bc0 = 0.  (Widening a mask's bounding box to the
    // left to get there was measured slower: up to a third more items.)
    if (!MASKED && (w & 15) == 0 && (bc0 & 15) == 0) {
        G.aligned = 1;
        const int other = (G.base_j ? -G.doff : G.doff) & 15;      // offset of the other side modulo 16 bytes
        G.ws = other >> 2;
        G.sb = (other & 3) << 3;
    }
    G.ipr = (G.c1 - G.c0 + 15) >> 4;               // items per row
    G.rcp = __frcp_rn((float)G.ipr);
    G.items = G.nrows * G.ipr;
    return G;
}

template <int WS>
__device__ __forceinline__ void k3_shift4(const uint4& a, const uint4& b, uint32_t sb, uint32_t (&out)[4]) {
    const uint32_t W[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) out[k] = __funnelshift_r(W[WS + k], W[WS + k + 1], sb);
}

// Load one item: the quantised bytes of both pixels of its 16 pairs (I4[k], J4[k]: pairs 4k..4k+3;
// bytes of pairs that do not exist are whatever lies there) and one bit per pair that exists (inside
// the image, and inside the mask when MASKED).  False if none does.
template <bool MASKED>
__device__ __forceinline__ bool k3_item16(const K3Group& Gp, const K3Geom& G, int item, uint32_t (&I4)[4],
                                          uint32_t (&J4)[4], uint32_t& pm) {
    const int r = (int)(((float)item + 0.5f) * G.rcp);
    const int c = G.c0 + 16 * (item - r * G.ipr);
    const int nv = min(16, G.c1 - c);
    const int oi = (G.r0 + r) * G.w + c, oj = oi + G.doff;
    pm = 0xffffu >> (16 - nv);
    if (!MASKED && G.aligned) {
        const int ob = G.base_j ? oj : oi, oo = G.base_j ? oi : oj;      // ob is a multiple of 16
        if (MASKED) {
            pm &= (uint32_t)reinterpret_cast<const uint16_t*>(Gp.mbits)[ob >> 4] & k3_bits16(Gp.mbits, oo);
            if (pm == 0u) return false;
        }
        const uint4 vb = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(Gp.q8) + ob);
        const uint4* po = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(Gp.q8) + (oo & ~15));
        const uint4 v0 = po[0], v1 = po[1];
        uint32_t B[4] = {vb.x, vb.y, vb.z, vb.w}, O[4];
        switch (G.ws) {
            case 0: k3_shift4<0>(v0, v1, (uint32_t)G.sb, O); break;
            case 1: k3_shift4<1>(v0, v1, (uint32_t)G.sb, O); break;
            case 2: k3_shift4<2>(v0, v1, (uint32_t)G.sb, O); break;
            default: k3_shift4<3>(v0, v1, (uint32_t)G.sb, O); break;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) { I4[k] = G.base_j ? O[k] : B[k]; J4[k] = G.base_j ? B[k] : O[k]; }
        return true;
    }
    if (MASKED) {
        pm &= k3_bits16(Gp.mbits, oi) & k3_bits16(Gp.mbits, oj);
        if (pm == 0u) return false;
    }
    const uint32_t* bi = Gp.q8 + (oi >> 2);
    const uint32_t* bj = Gp.q8 + (oj >> 2);
    const uint32_t si = (uint32_t)(oi & 3) << 3, sj = (uint32_t)(oj & 3) << 3;
    uint32_t wi[5], wj[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) { wi[k] = bi[k]; wj[k] = bj[k]; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        I4[k] = __funnelshift_r(wi[k], wi[k + 1], si);
        J4[k] = __funnelshift_r(wj[k], wj[k + 1], sj);
    }
    return true;
}

// Pair-stream sums of one group of 4 pairs; vm = 0xff per pair that exists.  Bytes of pairs that do not
// exist are zeroed, so they add nothing to the integer sums and exactly homtab[0] = 1.0 each to the
// homogeneity sum, which the epilogue subtracts again (16 per item minus the pair count).
__device__ __forceinline__ void k3_sums(const double* homtab, uint32_t I4, uint32_t J4, uint32_t vm, K3Acc& A) {
    I4 &= vm;
    J4 &= vm;
    A.si = __dp4a(I4, 0x01010101u, A.si);
    A.sj = __dp4a(J4, 0x01010101u, A.sj);
    A.sii = __dp4a(I4, I4, A.sii);
    A.sjj = __dp4a(J4, J4, A.sjj);
    A.sij = __dp4a(I4, J4, A.sij);
    A.sd += __vsadu4(I4, J4);
    const uint32_t D4 = __vabsdiffu4(I4, J4);
    A.hom += homtab[D4 & 0xffu];
    A.hom += homtab[(D4 >> 8) & 0xffu];
    A.hom += homtab[(D4 >> 16) & 0xffu];
    A.hom += homtab[D4 >> 24];
}

// ---- the bins ------------------------------------------------------------------------------------
// Bin (i, j) is a 16-bit counter: half j & 1 of word (i << 7 | ((j >> 1) ^ (i & 31) << 2)) -- the word
// index is swizzled with the low bits of i because neighbouring pixels have similar levels and the
// bank would otherwise depend on j alone (measured: 7.2 wavefronts per ATOMS without the swizzle).
// While the group does not own the table, every pair becomes a "hit" in a register:
//     hit = word << 16 | 1 << 8 * (j & 1)
// (a pair that does not exist -- row tail, outside the mask -- gets an increment of zero on whatever
// word the bytes lying there give), so the table phase is branch-free and four instructions per pair:
//     addr = hit >> 14;  inc = PRMT(hit) = 1 << 16 * (j & 1);  old = ATOMS.ADD [addr], inc;
//     sold = IDP.2A(old, hit) + sold       (= old count of that bin)
// sum_bins c^2 = 2 * sum(old) + M  (c^2 = sum_{k<c} (2k+1)), so the bins are never read back; clearing
// re-walks the hits.  (Merging equal hits of a warp with match.any first was measured 4x slower.)

// two hits from two 16-bit keys i << 8 | j in K; M = 0xffff per key whose pair exists
__device__ __forceinline__ void k3_hits2(uint32_t K, uint32_t M, uint32_t& ha, uint32_t& hb) {
    const uint32_t W = ((K >> 1) & 0x7fff7fffu) ^ ((K >> 6) & 0x007c007cu);       // swizzled words
    const uint32_t S = ((K & 0x00010001u) * 0xffu + 0x00010001u) & M;             // 1 << 8 * (j & 1), or 0
    ha = __byte_perm(S, W, 0x5410);
    hb = __byte_perm(S, W, 0x7632);
}
// the 16 hits of an item; pairs that do not exist keep whatever word their bytes give, with increment 0
__device__ __forceinline__ void k3_hits16(const uint32_t (&I4)[4], const uint32_t (&J4)[4], uint32_t pm, uint32_t (&h)[16]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t vm = k3_expand4(pm >> (4 * k));
        k3_hits2(__byte_perm(J4[k], I4[k], 0x5140), __byte_perm(vm, 0u, 0x1100), h[4 * k], h[4 * k + 1]);
        k3_hits2(__byte_perm(J4[k], I4[k], 0x7362), __byte_perm(vm, 0u, 0x3322), h[4 * k + 2], h[4 * k + 3]);
    }
}
// increments the bin of a hit and accumulates its old count
__device__ __forceinline__ void k3_hit(uint32_t hist_addr, uint32_t h, uint32_t& sold) {
    const uint32_t inc = __byte_perm(h, 0u, 0x4140);
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(hist_addr + (h >> 14)), "r"(inc) : "memory");
    sold = __dp2a_lo(old, h, sold);
}
__device__ __forceinline__ void k3_unhit(uint32_t hist_addr, uint32_t h) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(hist_addr + (h >> 14)), "r"(0u) : "memory");
}

// One direction's six properties from the exact integer sums.
__device__ __forceinline__ void k3_epilogue(const Params& P, double* out_row, uint32_t* status, int slot, int a,
                                            const K3AccS& A) {
    double* o = out_row + P.col_glcm + (slot * P.n_angles + a) * kNGlcm;
    const long long M = A.s[7];
    if (M == 0) {
        o[0] = 0.0; o[1] = 0.0; o[2] = 0.0; o[3] = 0.0; o[4] = 0.0; o[5] = 1.0;
        if (status) atomicOr(status, kStNoPairs);
        return;
    }
    // the walked pairs that do not exist added exactly 1.0 each to the homogeneity sum
    const unsigned long long hom_true = ((unsigned long long)A.hom_hi << 32 | A.hom_lo) - ((unsigned long long)((long long)A.np - M) << 40);
    const double Md = (double)M;
    const long long Si = A.s[0], Sj = A.s[1], Sii = A.s[2], Sjj = A.s[3], Sij = A.s[4];
    const long long vi = M * Sii - Si * Si, vj = M * Sjj - Sj * Sj, cov = M * Sij - Si * Sj;
    const double asmv = (double)(2ull * (unsigned long long)A.s[6] + (unsigned long long)M) / (Md * Md);
    o[0] = (double)(Sii + Sjj - 2 * Sij) / Md;
    o[1] = (double)A.s[5] / Md;
    o[2] = ((double)hom_true * 9.094947017729282e-13) / Md;
    o[3] = asmv;
    o[4] = sqrt(asmv);
    o[5] = (vi == 0 || vj == 0) ? 1.0 : (double)cov / (sqrt((double)vi) * sqrt((double)vj));
}

template <bool MASKED, bool DUMP, int NG>
__global__ void __launch_bounds__(kK3Threads, 1)
k3_glcm_kernel(const __grid_constant__ Params P, int max_pixels, int prefetch) {
    extern __shared__ __align__(16) unsigned char k3_smem_raw[];
    K3Smem& S = *reinterpret_cast<K3Smem*>(k3_smem_raw);
    constexpr int gthreads = kK3Threads / NG;
    constexpr int kCache = 256 / gthreads + (gthreads > 256 ? 1 : 0);   // items (16 pairs each) per thread held in registers
    const int tid = threadIdx.x, lane = tid & 31;
    const int g = tid / gthreads, gt = tid % gthreads, gw = gt >> 5;
    K3Group Gp;
    Gp.q8 = reinterpret_cast<uint32_t*>(k3_smem_raw + sizeof(K3Smem) + (size_t)g * k3_group_bytes(max_pixels, MASKED));
    Gp.mbits = Gp.q8 + k3_q8_words(max_pixels);
    // raw prefetch buffers lie behind the staging buffers of all groups
    unsigned char* raw = k3_smem_raw + sizeof(K3Smem) + (size_t)NG * k3_group_bytes(max_pixels, MASKED) +
                         (size_t)g * k3_raw_bytes(max_pixels, MASKED);
    const uint32_t hist_addr = smem_addr(S.hist);
    const int id_sync = 1 + g, id_mine = 1 + NG + g, id_next = 1 + NG + (g + 1) % NG;

    for (int k = tid; k < 32768; k += kK3Threads) S.hist[k] = 0u;
    for (int k = tid; k < 4 * 2 * kK3Park * kMaxAngles * (int)(sizeof(K3AccS) / 4); k += kK3Threads)
        reinterpret_cast<uint32_t*>(&S.acc[0][0][0][0])[k] = 0u;
    if (tid < 256) S.homtab[tid] = 1.0 / (1.0 + (double)(tid * tid));
    if (tid < 8) { S.box[tid >> 1][tid & 1][0] = 1 << 30; S.box[tid >> 1][tid & 1][1] = -1; S.box[tid >> 1][tid & 1][2] = 1 << 30; S.box[tid >> 1][tid & 1][3] = -1; }
    __syncthreads();
    // K1 (same stream, earlier launch) already wrote the tile maximum into the table when the basic
    // block is requested; then the max pass and its barrier are skipped.
    const bool k1_max = P.col_basic >= 0;

    // tiles of this CTA: blockIdx.x + k * gridDim.x; group g takes k = NG * j + g
    const uint32_t n_tiles = (uint32_t)P.n_tiles, first = blockIdx.x;
    const uint32_t mine = first < n_tiles ? (n_tiles - first + gridDim.x - 1) / gridDim.x : 0u;
    const uint32_t n_iter = (mine + NG - 1) / NG;          // every group runs the same number of rounds
    const uint32_t my_count = (mine + NG - 1 - g) / NG;
    const uint32_t t_step = NG * gridDim.x;
    uint32_t t = first + g * gridDim.x;                    // this round's tile
    if (g == NG - 1) bar_arrive(1 + NG, 2 * gthreads);     // the table starts out free for group 0
    // While a tile is under way the group prepares its next one: one lane (of the last warp, the least
    // loaded one) resolves it -- geometry, source pointers, K1's maximum (the load stays in flight until
    // publish() turns it into quantiser constants before the end-of-tile barrier) --, then every thread
    // starts its share of the cp.async copies of the raw pixels and mask bytes into the prefetch buffer.
    constexpr int gwarps = gthreads / 32;
    const bool scout = gw == gwarps - 1 && lane == 0;
    double vnext = 0.0;
    auto resolve = [&](uint32_t tile, int nb) {            // scout only
        const Tile T = resolve_tile(P, tile);
        K3TileInfo& I = S.info[g][nb];
        I.h = T.h; I.w = T.w; I.slot = (uint32_t)T.slot; I.row = tile / (uint32_t)P.c_out;
        I.px = T.px; I.mk = T.mk; I.n16 = (T.n * 2 + 15) >> 4; I.n8 = (T.n + 7) >> 3;
        if (k1_max) vnext = T.out_row[P.col_basic + kNBasic * T.slot + 10];
    };
    auto start_copies = [&](int nb) {                      // every thread of the group
        if (!prefetch) return;
        const K3TileInfo& I = S.info[g][nb];
        const unsigned char* src = reinterpret_cast<const unsigned char*>(I.px);
        for (int k = gt; k < I.n16; k += gthreads) cp_async16(raw + 16 * k, src + 16 * k);
        if (MASKED)
            for (int k = gt; k < I.n8; k += gthreads) cp_async8(raw + 2 * (size_t)max_pixels + 8 * k, I.mk + 8 * k);
    };
    auto publish = [&](int nb) {                           // scout only
        if (!k1_max) return;
        K3TileInfo& I = S.info[g][nb];
        const uint32_t vmax = (vnext == vnext) ? (uint32_t)vnext : 0u;   // NaN: empty mask, no pair exists anyway
        uint32_t mul = 0u, sh = 24u;
        const bool fast = k3_magic_fast(vmax, mul);
        if (!fast) k3_magic(vmax, mul, sh);
        I.mul = mul; I.sh = sh; I.fast = fast ? 1u : 0u;
    };
    // Epilogues (FP64 divisions and square roots, ~150 instructions per tile and direction) are parked: a
    // tile's sums stay in its parking slot, and once kK3Park tiles are parked the last warp of the group
    // finishes all of them at once, one (tile, direction) per lane.
    auto run_epilogues = [&](int bank, int count) {        // last warp of the group
        const int slot = lane / kMaxAngles, d = lane % kMaxAngles;
        if (gw == gwarps - 1 && slot < count && d < P.n_angles) {
            const uint32_t row = S.tileid[g][bank][slot][0];
            K3AccS& Acc = S.acc[g][bank][slot][d];
            k3_epilogue(P, P.out + (long long)row * P.row_stride, P.status ? P.status + row : nullptr,
                        (int)S.tileid[g][bank][slot][1], d, Acc);
#pragma unroll
            for (int k = 0; k < 8; ++k) Acc.s[k] = 0u;
            Acc.hom_lo = 0u; Acc.hom_hi = 0u;
            Acc.np = 0u;
        }
    };
    static_assert(kK3Park * kMaxAngles <= 32, "one lane per parked (tile, direction)");
    int parked = 0, pbank = 0;                             // parking slot and bank of the current tile
    if (scout && my_count) resolve(t, 0);
    bar_sync(id_sync, gthreads);
    if (my_count) start_copies(0);
    cp_async_wait_all();
    if (scout && my_count) publish(0);
    bar_sync(id_sync, gthreads);

    for (uint32_t j = 0; j < n_iter; ++j, t += t_step) {
        const bool active = j < my_count;
        const int buf = (int)(j & 1u);
        int tw = 0, th = 0;
        if (scout && j + 1 < my_count) resolve(t + t_step, buf ^ 1);
        if (active) {
            const K3TileInfo& I = S.info[g][buf];
            tw = I.w; th = I.h;
            if (gt == 0) { S.tileid[g][pbank][parked][0] = I.row; S.tileid[g][pbank][parked][1] = I.slot; }
            const int tn = th * tw;
            // pixels and mask bytes: from the prefetch buffer, or straight from global memory
            const uint16_t* pxs = reinterpret_cast<const uint16_t*>(raw);
            const uint8_t* mks = raw + 2 * (size_t)max_pixels;
            if (!prefetch) { const Tile T = resolve_tile(P, t); pxs = T.px; mks = T.mk; }
            const uint4* px4 = reinterpret_cast<const uint4*>(pxs);
            const uint2* mk2 = reinterpret_cast<const uint2*>(mks);
            const int nfull = tn >> 3, rem = tn & 7;
            uint8_t* mbytes = reinterpret_cast<uint8_t*>(Gp.mbits);

            // ---- 1. tile maximum (over the mask when masked); stage the mask bits and their bounding box ----
            uint32_t mx2 = 0u;
            int brmin = 1 << 30, brmax = -1, bcmin = 1 << 30, bcmax = -1;
            const float rtw = __frcp_rn((float)tw);
            for (int idx = gt; idx < nfull && (MASKED || !k1_max); idx += gthreads) {
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (!k1_max) v = px4[idx];
                if (MASKED) {
                    const uint2 m = mk2[idx];
                    const uint32_t c0 = __vcmpne4(m.x, 0u), c1 = __vcmpne4(m.y, 0u);
                    // 8 mask bytes -> 8 bits (byte k -> bit k)
                    const uint32_t b0 = ((c0 & 0x01010101u) * 0x01020408u) >> 24;
                    const uint32_t b1 = ((c1 & 0x01010101u) * 0x01020408u) >> 24;
                    const uint32_t bits8 = (b0 & 0xfu) | ((b1 & 0xfu) << 4);
                    mbytes[idx] = (uint8_t)bits8;
                    if (bits8) {
                        const int p0 = 8 * idx, ra = (int)(((float)p0 + 0.5f) * rtw), ca = p0 - ra * tw;   // exact: p0 < 2^20
                        if (ca + 7 < tw) {                 // the 8 pixels lie in one row
                            brmin = min(brmin, ra); brmax = max(brmax, ra);
                            bcmin = min(bcmin, ca + __ffs(bits8) - 1); bcmax = max(bcmax, ca + 31 - __clz(bits8));
                        } else {                           // straddles rows: be conservative
                            brmin = min(brmin, ra); brmax = max(brmax, (p0 + 7) / tw);
                            bcmin = 0; bcmax = tw - 1;
                        }
                    }
                    v.x &= __byte_perm(c0, 0u, 0x1100); v.y &= __byte_perm(c0, 0u, 0x3322);
                    v.z &= __byte_perm(c1, 0u, 0x1100); v.w &= __byte_perm(c1, 0u, 0x3322);
                }
                mx2 = __vmaxu2(mx2, __vmaxu2(__vmaxu2(v.x, v.y), __vmaxu2(v.z, v.w)));
            }
            if (gt == 0 && rem) {                          // tail pixels (< 8): one thread, in order
                uint32_t bits = 0u;
                for (int k = 0; k < rem; ++k) {
                    const int i = nfull * 8 + k;
                    const bool ok = !MASKED || mks[i] != 0;
                    if (ok) {
                        bits |= 1u << k;
                        if (!k1_max) mx2 = __vmaxu2(mx2, (uint32_t)pxs[i]);
                        const int ra = i / tw, ca = i - ra * tw;
                        brmin = min(brmin, ra); brmax = max(brmax, ra); bcmin = min(bcmin, ca); bcmax = max(bcmax, ca);
                    }
                }
                if (MASKED) mbytes[nfull] = (uint8_t)bits;
            }
            if (MASKED) {
                brmin = __reduce_min_sync(0xffffffffu, brmin); brmax = __reduce_max_sync(0xffffffffu, brmax);
                bcmin = __reduce_min_sync(0xffffffffu, bcmin); bcmax = __reduce_max_sync(0xffffffffu, bcmax);
                if (lane == 0 && brmax >= 0) {
                    atomicMin(&S.box[g][buf][0], brmin); atomicMax(&S.box[g][buf][1], brmax);
                    atomicMin(&S.box[g][buf][2], bcmin); atomicMax(&S.box[g][buf][3], bcmax);
                }
            }
            uint32_t mul = I.mul, sh = I.sh;
            bool fast = I.fast != 0u;
            if (!k1_max) {
                const uint32_t wm = __reduce_max_sync(0xffffffffu, max(mx2 & 0xffffu, mx2 >> 16));
                if (lane == 0) S.wmax[g][gw] = wm;
                bar_sync(id_sync, gthreads);
                uint32_t vmax = lane < gthreads / 32 ? S.wmax[g][lane] : 0u;
                vmax = __reduce_max_sync(0xffffffffu, vmax);
                uint32_t f = 0u;
                if (lane == 0) { f = k3_magic_fast(vmax, mul) ? 1u : 0u; if (!f) k3_magic(vmax, mul, sh); }
                mul = __shfl_sync(0xffffffffu, mul, 0);
                sh = __shfl_sync(0xffffffffu, sh, 0);
                fast = __shfl_sync(0xffffffffu, f, 0) != 0u;
            }

            // ---- 2. quantise to 8 bits into shared memory (out-of-mask pixels may exceed the maximum: their
            //         bytes are never part of a pair, only the low byte is kept) ----
            if (fast) {
                for (int idx = gt; idx < nfull; idx += gthreads) {
                    const uint4 v = px4[idx];
                    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
                    uint32_t q[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        q[k] = __byte_perm(__umulhi(w4[k] & 0xffffu, mul), __umulhi(w4[k] >> 16, mul), 0x0040);
                    *reinterpret_cast<uint2*>(Gp.q8 + 2 * idx) = make_uint2(__byte_perm(q[0], q[1], 0x5410), __byte_perm(q[2], q[3], 0x5410));
                }
            } else {
                for (int idx = gt; idx < nfull; idx += gthreads) {
                    const uint4 v = px4[idx];
                    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
                    uint32_t q[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        q[k] = __byte_perm(k3_quant(w4[k] & 0xffffu, mul, sh), k3_quant(w4[k] >> 16, mul, sh), 0x0040);
                    *reinterpret_cast<uint2*>(Gp.q8 + 2 * idx) = make_uint2(__byte_perm(q[0], q[1], 0x5410), __byte_perm(q[2], q[3], 0x5410));
                }
            }
            if (gt < rem) {
                const int i = nfull * 8 + gt;
                const uint32_t qv = fast ? __umulhi((uint32_t)pxs[i], mul) : k3_quant(pxs[i], mul, sh);
                reinterpret_cast<uint8_t*>(Gp.q8)[i] = (uint8_t)qv;
            }
        }
        bar_sync(id_sync, gthreads);                       // this tile staged
        if (j + 1 < my_count) start_copies(buf ^ 1);       // the raw buffer is free again
        // box of pixels that can take part in a pair: the tile, or the mask's bounding box
        int bx[4] = {0, th - 1, 0, tw - 1};
        if (MASKED && active) { bx[0] = S.box[g][buf][0]; bx[1] = S.box[g][buf][1]; bx[2] = S.box[g][buf][2]; bx[3] = S.box[g][buf][3]; }

#pragma unroll
        for (int a = 0; a < kMaxAngles; ++a) {
            if (a >= P.n_angles) break;
            // ---- off the table: this direction's first items become hits in registers ----
            const K3Geom G = k3_geom<MASKED>(tw, P.dr[a], P.dc[a], bx[0], bx[1], bx[2], bx[3]);
            uint32_t hit[kCache][16];
            uint32_t sold = 0u, valid = 0u, np = 0u;
            K3Acc A = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0.0};
            K3AccS& Acc = S.acc[g][pbank][parked][a];
            auto sums16 = [&](const uint32_t (&I4)[4], const uint32_t (&J4)[4], uint32_t pm) {
#pragma unroll
                for (int k = 0; k < 4; ++k) k3_sums(S.homtab, I4[k], J4[k], k3_expand4(pm >> (4 * k)), A);
                A.m += __popc(pm);
                np += 16u;
            };
            // warp totals of the pair-stream sums into the group's accumulators
            auto flush_sums = [&]() {
                if (!__any_sync(0xffffffffu, np != 0u)) return;
                const uint32_t r0 = __reduce_add_sync(0xffffffffu, A.si);
                const uint32_t r1 = __reduce_add_sync(0xffffffffu, A.sj);
                const uint32_t r2 = __reduce_add_sync(0xffffffffu, A.sii);
                const uint32_t r3 = __reduce_add_sync(0xffffffffu, A.sjj);
                const uint32_t r4 = __reduce_add_sync(0xffffffffu, A.sij);
                const uint32_t r5 = __reduce_add_sync(0xffffffffu, A.sd);
                const uint32_t r7 = __reduce_add_sync(0xffffffffu, A.m);
                const uint32_t r8 = __reduce_add_sync(0xffffffffu, np);
                // per-thread double sums (fixed order) are rounded to 2^-40 fixed point: the sums over lanes
                // and warps are integer and do not depend on their order
                const unsigned long long hf = warp_sum_redux((unsigned long long)__double2ll_rn(A.hom * 1099511627776.0));
                if (lane == 0) {
                    atomicAdd(&Acc.s[0], r0); atomicAdd(&Acc.s[1], r1); atomicAdd(&Acc.s[2], r2); atomicAdd(&Acc.s[3], r3);
                    atomicAdd(&Acc.s[4], r4); atomicAdd(&Acc.s[5], r5); atomicAdd(&Acc.s[7], r7);
                    atomicAdd(&Acc.np, r8);
                    const uint32_t lo = (uint32_t)hf, old = atomicAdd(&Acc.hom_lo, lo);
                    atomicAdd(&Acc.hom_hi, (uint32_t)(hf >> 32) + (old + lo < old ? 1u : 0u));
                }
                A = K3Acc{0u, 0u, 0u, 0u, 0u, 0u, 0u, 0.0};
                np = 0u;
            };
#pragma unroll
            for (int i = 0; i < kCache; ++i) {
                const int item = gt + i * gthreads;
                uint32_t I4[4], J4[4], pm;
                if (item < G.items && k3_item16<MASKED>(Gp, G, item, I4, J4, pm)) {
                    sums16(I4, J4, pm);
                    k3_hits16(I4, J4, pm, hit[i]);
                    valid |= 1u << i;
                }
            }
            flush_sums();
            bar_sync(id_mine, 2 * gthreads);               // ---- table owned by this group ----
#pragma unroll
            for (int i = 0; i < kCache; ++i)
                if (valid & (1u << i)) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) k3_hit(hist_addr, hit[i][k], sold);
                }
            for (int item = gt + kCache * gthreads; item < G.items; item += gthreads) {
                uint32_t I4[4], J4[4], pm, h16[16];
                if (k3_item16<MASKED>(Gp, G, item, I4, J4, pm)) {
                    sums16(I4, J4, pm);
                    k3_hits16(I4, J4, pm, h16);
#pragma unroll
                    for (int k = 0; k < 16; ++k) k3_hit(hist_addr, h16[k], sold);
                }
            }
            bar_sync(id_sync, gthreads);                   // bins of this direction complete
            if (DUMP) {
                if (active) {
                    uint32_t* dst = P.counts + ((long long)t * P.n_angles + a) * 65536ll;
                    for (int k = gt; k < 32768; k += gthreads) {
                        const uint32_t wv = S.hist[k];
                        const int nat = k ^ (((k >> 7) & 31) << 2);          // undo the bank swizzle
                        reinterpret_cast<uint2*>(dst)[nat] = make_uint2(wv & 0xffffu, wv >> 16);
                    }
                }
                bar_sync(id_sync, gthreads);
            }
#pragma unroll
            for (int i = 0; i < kCache; ++i)
                if (valid & (1u << i)) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) k3_unhit(hist_addr, hit[i][k]);
                }
            for (int item = gt + kCache * gthreads; item < G.items; item += gthreads) {
                uint32_t I4[4], J4[4], pm, h16[16];
                if (k3_item16<MASKED>(Gp, G, item, I4, J4, pm)) {
                    k3_hits16(I4, J4, pm, h16);
#pragma unroll
                    for (int k = 0; k < 16; ++k) k3_unhit(hist_addr, h16[k]);
                }
            }
            // ---- hand the clean table on (the very last hand-over has no taker) ----
            if (!(g == NG - 1 && j + 1 == n_iter && a + 1 == P.n_angles)) bar_arrive(id_next, 2 * gthreads);
            if (G.items > kCache * gthreads) flush_sums();  // items beyond the register cache
            const uint32_t so = __reduce_add_sync(0xffffffffu, sold);
            if (lane == 0 && so) atomicAdd(&Acc.s[6], so);
        }
        cp_async_wait_all();
        if (scout && j + 1 < my_count) publish(buf ^ 1);
        bar_sync(id_sync, gthreads);                       // sums final, staging buffers free, next tile prepared
        if (active && ++parked == kK3Park) { run_epilogues(pbank, kK3Park); parked = 0; pbank ^= 1; }
        if (MASKED && active && gt == 0) {
            S.box[g][buf][0] = 1 << 30; S.box[g][buf][1] = -1; S.box[g][buf][2] = 1 << 30; S.box[g][buf][3] = -1;
        }
    }
    run_epilogues(pbank, parked);
}

}  // namespace ring
}  // namespace imfeat
