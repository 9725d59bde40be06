// k3_glcm.cuh -- K3: gray-level quantisation + GLCM + Haralick properties.
//
// Replaces, per channel:
//   (x / x.max()) * 255 -> uint8                      NB:293-295
//   greycomatrix(q, [5], [0], levels=256)              NB:298   (not symmetric, not normed)
//   greycoprops x6 (contrast .. correlation)           NB:301-306
//
// * quantiser: floor(255*x/max) with an exact multiply-shift reciprocal (bit-identical to the
//   notebook's float64 expression for every uint16 pair; tests/test_oracle_cpu.py).
// * one small CTA per tile, three CTAs per SM.  Each CTA owns the 256x256 bins of ONE (tile, direction)
//   at a time as 8-bit counters (four per 32-bit word, 64 KB of shared memory), so no table is ever
//   shared between tiles and nothing is handed over: a direction is
//       build   every pair adds 1 to its counter (fire-and-forget shared-memory atomics)
//       barrier
//       clear   every pair exchanges its word with 0; the first to arrive gets the four counts of the
//               word and adds their squares (one DP4A) -- sum_bins c^2 (ASM) needs no pass over bins,
//               and the table is clean again for the next direction
//       barrier
//   The sum of the counts that come back must equal the number of pairs; if it does not, an 8-bit
//   counter wrapped (more than 255 pairs in one bin: flat backgrounds), and that direction alone is
//   done again with 16-bit counters in two passes over half the key space each.  The result is exact
//   either way.
// * contrast, dissimilarity, homogeneity and correlation are pair-stream sums (no bins needed), four
//   pairs per IDP.4A / VABSDIFF4; homogeneity through a 256-entry FP64 table.
// * the kernel leaves the raw integer sums of a (tile, direction) -- twelve 32-bit words -- in the six
//   output doubles of that direction; k3_finalize_kernel turns them into the six properties in place
//   (FP64 divisions and square roots, one thread per record, off the table kernel's critical path).
#pragma once
#include "common.cuh"

namespace imfeat {

constexpr int kK3Rec = 12;           // raw 32-bit words per (tile, direction) = the six output doubles
enum { kR_si = 0, kR_sj, kR_sii, kR_sjj, kR_sij, kR_sd, kR_sq, kR_m, kR_homlo, kR_homhi, kR_np, kR_cnt };
constexpr int kK3MaxWarps = 8;

// staging buffers behind K3Smem: quantised pixels (one byte each, row-major, + slack for the 16-byte
// item reads) and mask bits (masked variant only)
__host__ __device__ inline int k3_q8_words(int max_pixels) { return (max_pixels / 4 + 8 + 3) & ~3; }
__host__ __device__ inline int k3_mb_words(int max_pixels, bool masked) { return masked ? ((max_pixels / 32 + 2 + 3) & ~3) : 0; }
// What a launch of the two kernels can hold per tile: q8 bytes of quantised rows, mask bits for mb pixels.  A batch
// of large strides is worked off in two tiers: first with room for about a third of a full tile's rows -- enough
// for the bounding box of a sparse mask, and three times as many warps / CTAs per SM --, then the tiles whose
// box did not fit (the front kernel lists them) with room for everything.
struct K3Cap { int q8, mb; };
struct K3Tier {
    const uint32_t* wl_in;       // tiles of this launch (second tier), or NULL: tile_base + 0 .. n_local - 1
    const uint32_t* n_in;        // their number (device), with wl_in
    uint32_t* wl_out;            // first tier: where tiles that do not fit are listed (NULL: every tile fits)
    uint32_t* n_out;
    unsigned int* counter;       // the launch's dynamic tile counter
};
constexpr uint32_t kK3Skip = 16u;      // record length of a tile left to the second tier (a real record is longer)

// An item is a run of up to 4*NG horizontally consecutive pairs (NG groups of 4) of one row: the words
// of both pixel runs are loaded once and funnel-shifted into place.  Unmasked tiles use 16-pair items;
// when the row length is a multiple of 16 the runs of one side (I for dc >= 0, J for dc < 0) start at
// multiples of 16 bytes: that side is one conflict-free 128-bit load, the other side two.  Masked tiles
// use 8-pair items (the bounding box of a mask is narrow: shorter items waste fewer lanes).
struct alignas(16) K3Geom {
    int nrows, r0, c0, c1, ipr, items, w, doff;
    float rcp;                                     // 1 / ipr: row = floor((item + 0.5) * rcp), exact for item < 2^20
    int aligned, base_j, ws, sb;                   // aligned path: which side is 16-byte aligned; word / bit shift of the other
                                                   // masked tiles: sb = byte offset of pixel (r0, 0) in the staged rows
    int pad[3];
};
static_assert(sizeof(K3Geom) == 64, "four 128-bit shared-memory loads");

// What the front kernel (K3a) leaves for the bins kernel (K3b) per tile: this header, then the mask bits (all
// rows), then the quantised pixels of the rows the mask's bounding box spans -- one contiguous record of the
// scratch buffer, fetched with one bulk copy of `len` bytes (a mask that covers a tenth of a 128x128 tile makes
// a 5 KB record, not 19 KB; the lengths are kept in a table behind the records, because the bins kernel must
// know them before it fetches).
struct alignas(16) K3Hdr {
    K3Geom geom[kMaxAngles];                       // pad[0] = number of pairs of that direction
    uint32_t* rec;                                 // the tile's first record (direction 0) in the output row
    uint32_t tile, len;                            // len: bytes of this record that are in use (multiple of 16)
};
static_assert(sizeof(K3Hdr) == 272, "multiple of 16 bytes");
__host__ __device__ inline size_t k3_rec_bytes(K3Cap cap, bool masked) {
    return sizeof(K3Hdr) + 4 * (size_t)(k3_q8_words(cap.q8) + k3_mb_words(cap.mb, masked));
}

struct alignas(16) K3Smem {                     // behind the table (64 or 32 KB), before the two record buffers
    uint32_t part[2][kK3MaxWarps][2];           // per direction parity and warp: sum of squares, counts taken back
    unsigned long long mbar[2];                 // completion of the bulk copy into record buffer 0 / 1
    uint32_t tq[4];                             // local tile indices drawn from the global counter
    uint32_t tlen[4];                           // ... and the lengths of their records
    uint32_t slow[4];                           // fallback: sum of squares, [1] dense count check of the dump, [2] counts taken back
};
struct K3Group {                               // where the quantised tile and its mask bits live
    uint32_t* q8;                              // quantised pixels (bytes) + slack for unaligned reads
    uint32_t* mbits;                           // one bit per pixel: inside the mask (masked variant)
#ifdef IMFEAT_CHECKS
    int q8_words, mb_words;                    // their sizes, for the bounds checks
#endif
};
__host__ __device__ inline size_t k3_smem_bytes(K3Cap max_pixels, bool masked, int table_kb) {
    // two record buffers: the next tile lands while this one is worked off
    return (size_t)table_kb * 1024 + sizeof(K3Smem) + 2 * k3_rec_bytes(max_pixels, masked);
}
// front kernel: per warp one record being assembled + the four output records
constexpr int kK3aWarps = 4;                    // warps per CTA of the front kernel
__host__ __device__ inline size_t k3a_warp_bytes(K3Cap max_pixels, bool masked) {
    return k3_rec_bytes(max_pixels, masked) + 16 * ((kMaxAngles * kK3Rec * 4 + 15) / 16);
}

struct K3Acc {
    uint32_t si, sj, sii, sjj, sij, sd, m;
    double hom0, hom1;
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

// consecutive mask bits starting at bit offset off (the caller masks what it needs)
__device__ __forceinline__ uint32_t k3_bits(const uint32_t* b, int off) {
    const int w = off >> 5;
    return __funnelshift_r(b[w], b[w + 1], off & 31);
}

// Row pitch (bytes) of the quantised tile in shared memory.  Masked tiles whose rows are a multiple of 8 pixels get
// an ODD number of words per row, and the lanes of a warp take items of consecutive ROWS (same column block): the
// 32 word loads of a warp then hit 32 different banks.  With the compact pitch of a 64-pixel row (16 words) and
// row-major items, rows two apart fell on the same banks: 3.4 wavefronts per load instead of 1 (ncu, source page),
// a third of all shared-memory traffic of the front kernel and more than half of the bins kernel's loads.
template <bool MASKED>
__host__ __device__ __forceinline__ int k3_pitch(int w) {
    return (MASKED && w >= 32 && (w & 7) == 0) ? (((w >> 2) | 1) << 2) : w;
}
// bytes of quantised pixels a batch of stride hs x ws can need (pitch <= w + 4)
__host__ __device__ inline int k3_q8_capacity(int hs, int ws, bool masked) { return ((masked ? hs * (ws + 4) : hs * ws) + 7) & ~7; }

// Pairs (r, c) -> (r + dr, c + dc) with both pixels inside the box rows [br0, br1], columns
// [bc0, bc1] (the whole tile, or the bounding box of the mask: pairs outside it cannot exist).
template <bool MASKED, int NG>
__device__ __forceinline__ K3Geom k3_geom(int w, int dr, int dc, int br0, int br1, int bc0, int bc1, int qbias) {
    K3Geom G;
    const int pitch = k3_pitch<MASKED>(w);
    G.r0 = br0;
    G.nrows = br1 - dr - br0 + 1;                  // dr >= 0 for all supported directions
    G.c0 = bc0 + (dc < 0 ? -dc : 0);
    G.c1 = bc1 + 1 - (dc > 0 ? dc : 0);
    G.w = pitch;                                   // quantised pixels: row pitch and pair offset in bytes
    G.doff = dr * pitch + dc;
    G.aligned = 0; G.base_j = dc < 0; G.ws = 0; G.sb = MASKED ? br0 * pitch - qbias : 0;
    G.pad[0] = 0; G.pad[1] = w; G.pad[2] = dr * w + dc;      // mask bits stay compact: their row pitch and pair offset
    if (G.nrows <= 0 || G.c1 <= G.c0) { G.items = 0; G.ipr = 1; G.rcp = 1.0f; G.nrows = 0; return G; }
    // unmasked tiles only: the aligned side starts at column bc0 = 0.  (Widening a mask's bounding box to the
    // left to get there was measured slower: up to a third more items.)
    if (!MASKED && NG == 4 && (w & 15) == 0 && (bc0 & 15) == 0) {
        G.aligned = 1;
        const int other = (G.base_j ? -G.doff : G.doff) & 15;      // offset of the other side modulo 16 bytes
        G.ws = other >> 2;
        G.sb = (other & 3) << 3;
    }
    G.ipr = (G.c1 - G.c0 + 4 * NG - 1) / (4 * NG); // items per row
    G.rcp = __frcp_rn((float)(MASKED ? G.nrows : G.ipr));     // masked: items run down the rows first (see k3_pitch)
    G.items = G.nrows * G.ipr;
    return G;
}

template <int WS>
__device__ __forceinline__ void k3_shift4(const uint4& a, const uint4& b, uint32_t sb, uint32_t (&out)[4]) {
    const uint32_t W[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) out[k] = __funnelshift_r(W[WS + k], W[WS + k + 1], sb);
}

// Load one item: the quantised bytes of both pixels of its 4*NG pairs (I4[k], J4[k]: pairs 4k..4k+3;
// bytes of pairs that do not exist are whatever lies there) and one bit per pair that exists (inside
// the image, and inside the mask when MASKED).  False if none does.
template <bool MASKED, int NG>
__device__ __forceinline__ bool k3_item(const K3Group& Gp, const K3Geom& G, int item, uint32_t (&I4)[NG],
                                        uint32_t (&J4)[NG], uint32_t& pm) {
    constexpr int NP = 4 * NG;
    int r, c;
    if (MASKED) {                                  // column-major: consecutive items = consecutive rows
        const int k = (int)(((float)item + 0.5f) * G.rcp);
        r = item - k * G.nrows;
        c = G.c0 + NP * k;
    } else {
        r = (int)(((float)item + 0.5f) * G.rcp);
        c = G.c0 + NP * (item - r * G.ipr);
    }
    const int nv = min(NP, G.c1 - c);
    const int oi = MASKED ? G.sb + r * G.w + c : (G.r0 + r) * G.w + c, oj = oi + G.doff;
    pm = ((1u << NP) - 1u) >> (NP - nv);
    if (!MASKED && NG == 4 && G.aligned) {
        const int ob = G.base_j ? oj : oi, oo = G.base_j ? oi : oj;      // ob is a multiple of 16
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
        for (int k = 0; k < NG; ++k) { I4[k] = G.base_j ? O[k & 3] : B[k & 3]; J4[k] = G.base_j ? B[k & 3] : O[k & 3]; }
        return true;
    }
    if (MASKED) {
        const int mi = (G.r0 + r) * G.pad[1] + c;
        IMFEAT_CHECK(mi >= 0 && mi + G.pad[2] >= 0 && (mi >> 5) + 1 < Gp.mb_words && ((mi + G.pad[2]) >> 5) + 1 < Gp.mb_words);
        pm &= k3_bits(Gp.mbits, mi) & k3_bits(Gp.mbits, mi + G.pad[2]);
        if (pm == 0u) return false;
    }
    IMFEAT_CHECK(oi >= 0 && oj >= 0 && (oi >> 2) + NG < Gp.q8_words && (oj >> 2) + NG < Gp.q8_words);
    const uint32_t* bi = Gp.q8 + (oi >> 2);
    const uint32_t* bj = Gp.q8 + (oj >> 2);
    const uint32_t si = (uint32_t)(oi & 3) << 3, sj = (uint32_t)(oj & 3) << 3;
    uint32_t wi[NG + 1], wj[NG + 1];
#pragma unroll
    for (int k = 0; k < NG + 1; ++k) { wi[k] = bi[k]; wj[k] = bj[k]; }
#pragma unroll
    for (int k = 0; k < NG; ++k) {
        I4[k] = __funnelshift_r(wi[k], wi[k + 1], si);
        J4[k] = __funnelshift_r(wj[k], wj[k + 1], sj);
    }
    return true;
}

// Pair-stream sums of one group of 4 pairs; vm = 0xff per pair that exists.  Bytes of pairs that do not
// exist are zeroed, so they add nothing to the integer sums and exactly homtab[0] = 1.0 each to the
// homogeneity sum, which the finalize kernel subtracts again (walked pairs minus the pair count).
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
#ifdef IMFEAT_EXP_NOHOM                                    // measurement only (wrong results): what the look-ups cost at most
    A.hom0 += (double)D4; return;
#endif
    A.hom0 += homtab[D4 & 0xffu];                  // two chains: a DADD waits several cycles for the one before it
    A.hom1 += homtab[(D4 >> 8) & 0xffu];
    A.hom0 += homtab[(D4 >> 16) & 0xffu];
    A.hom1 += homtab[D4 >> 24];
}

// ---- the bins ------------------------------------------------------------------------------------
// The counters of one (tile, direction) live in a table of TB = 64 KB (unmasked tiles) or 32 KB (masked
// tiles, whose bins rarely hold more than a handful of pairs: five CTAs fit an SM instead of three).  The
// table is used in one of these modes; the pairs that do not take part in a pass add 0:
//   MODE 0      64 KB  8-bit counters, all pairs      bin (i, j): byte j >> 6 of word i << 6 | x, x = (j & 63) ^ (i & 15) << 2
//   MODE 1, 2   64 KB  16-bit counters, i even / odd  bin (i >> 1, j): half j & 1 of word (i >> 1) << 7 | j >> 1
//   MODE 3      32 KB  4-bit counters, all pairs      bin (i, j): nibble j >> 5 of word i << 5 | x, x = (j & 31) ^ (i & 7) << 2
//   MODE 4, 5   32 KB  8-bit counters, i even / odd   bin (i >> 1, j): byte j >> 6 of word (i >> 1) << 6 | x, x = (j & 63) ^ (i >> 1 & 15) << 2
//   MODE 6..9   32 KB  16-bit counters, i & 3 = 0..3  bin (i >> 2, j): half j & 1 of word (i >> 2) << 7 | j >> 1
// In the packed modes the bins of a word are 64 (32) gray levels apart, so the similar levels of neighbouring
// pixels land in different words, and the word index is swizzled with i because the bank would otherwise depend
// on j alone (simulated on the synthetic tiles: 3.9 wavefronts per warp-wide atomic, against 6.3 for the plain
// layout and 3.5 for uniformly random words; measured: 3.96).
// k3_pack4 works on four pairs at once: R / X = high / low byte of the word's byte offset (shifted left by one
// in the 32 KB modes so that both fit a byte), H = shift of the increment, E = 1 per pair that takes part.
template <int MODE> struct K3Mode {
    static constexpr int kTableKB = MODE <= 2 ? 64 : 32;
    static constexpr int kAddrShift = MODE <= 2 ? 16 : 17;
    static constexpr int kBits = (MODE == 0 || MODE == 4 || MODE == 5) ? 8 : (MODE == 3 ? 4 : 16);
};
template <int MODE>
__device__ __forceinline__ void k3_pack4(uint32_t I4, uint32_t J4, uint32_t V1, uint32_t& X, uint32_t& R, uint32_t& H, uint32_t& E) {
    if (MODE == 0) {
        X = ((J4 << 2) & 0xfcfcfcfcu) ^ ((I4 << 4) & 0xf0f0f0f0u);
        R = I4;
        H = (J4 >> 3) & 0x18181818u;
        E = V1;
    } else if (MODE <= 2) {                        // byte offset i7..i1 j7..j1 00: the top bit of j goes into the row byte
        X = (J4 << 1) & 0xfcfcfcfcu;
        R = (I4 & 0xfefefefeu) | ((J4 >> 7) & 0x01010101u);
        H = (J4 << 4) & 0x10101010u;
        E = V1 & (MODE == 1 ? ~I4 : I4);           // V1 has only bit 0 of each byte
    } else if (MODE == 3) {                        // twice the byte offset: i7..i0 x4..x0 000
        X = ((J4 << 3) & 0xf8f8f8f8u) ^ ((I4 << 5) & 0xe0e0e0e0u);
        R = I4;
        H = (J4 >> 3) & 0x1c1c1c1cu;               // 4 * (j >> 5)
        E = V1;
    } else if (MODE <= 5) {                        // twice the byte offset: i7..i1 x5..x0 000
        const uint32_t XS = (J4 & 0x3f3f3f3fu) ^ ((I4 << 1) & 0x3c3c3c3cu);
        X = (XS << 3) & 0xf8f8f8f8u;
        R = (I4 & 0xfefefefeu) | ((XS >> 5) & 0x01010101u);
        H = (J4 >> 3) & 0x18181818u;
        E = V1 & (MODE == 4 ? ~I4 : I4);
    } else {                                       // twice the byte offset: i7..i2 j7..j1 000
        X = (J4 << 2) & 0xf8f8f8f8u;
        R = (I4 & 0xfcfcfcfcu) | ((J4 >> 6) & 0x03030303u);
        H = (J4 << 4) & 0x10101010u;
        const uint32_t t = I4 ^ ((uint32_t)(MODE - 6) * 0x01010101u);
        E = V1 & ~(t | (t >> 1));                  // the two low bits of i match
    }
}
// shared-window address of the word of pair K (0..3) of a group
template <int K, int SHIFT>
__device__ __forceinline__ uint32_t k3_addr(uint32_t hist_addr, uint32_t X, uint32_t R) {
    constexpr uint32_t sel = ((4u + K) << 12) | ((uint32_t)K << 8) | ((4u + K) << 4) | (uint32_t)K;
    return hist_addr + (__byte_perm(X, R, sel) >> SHIFT);        // R << 8 | X, halved in the 32 KB modes
}
// increment of pair K: 1 << shift if the pair takes part (E: byte K = 1), else 0 -- the add then
// changes nothing, on whatever word the bytes lying there give, and the build phase is branch-free
// (ptxas turns a predicated shared-memory atomic into a branch around it)
template <int K>
__device__ __forceinline__ uint32_t k3_inc(uint32_t H, uint32_t E) {
    return __funnelshift_l(0u, __byte_perm(E, 0u, 0x4440 + K), H >> (8 * K));   // only the low 5 bits of the amount count
}
__device__ __forceinline__ void k3_red(uint32_t addr, uint32_t inc) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(inc) : "memory");
}
// take the counts of a word, leaving it clean; squares and counts of what came back
template <int MODE>
__device__ __forceinline__ void k3_take(uint32_t addr, uint32_t& sq, uint32_t& cnt) {
    uint32_t old;
    asm volatile("atom.shared.exch.b32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(0u) : "memory");
    if (K3Mode<MODE>::kBits == 8) {
        sq = __dp4a(old, old, sq);
        cnt = __dp4a(old, 0x01010101u, cnt);
    } else if (K3Mode<MODE>::kBits == 4) {
        const uint32_t lo = old & 0x0f0f0f0fu, hi = (old >> 4) & 0x0f0f0f0fu;
        sq = __dp4a(hi, hi, __dp4a(lo, lo, sq));
        cnt = __dp4a(lo + hi, 0x01010101u, cnt);
    } else {
        const uint32_t c0 = old & 0xffffu, c1 = old >> 16;
        sq += c0 * c0 + c1 * c1;
        cnt += c0 + c1;
    }
}
template <int MODE>
__device__ __forceinline__ void k3_build4(uint32_t hist_addr, uint32_t I4, uint32_t J4, uint32_t V1, uint32_t (&keep)[4]) {
    constexpr int SH = K3Mode<MODE>::kAddrShift;
    uint32_t X, R, H, E;
    k3_pack4<MODE>(I4, J4, V1, X, R, H, E);
    keep[0] = k3_addr<0, SH>(hist_addr, X, R); keep[1] = k3_addr<1, SH>(hist_addr, X, R);
    keep[2] = k3_addr<2, SH>(hist_addr, X, R); keep[3] = k3_addr<3, SH>(hist_addr, X, R);
    IMFEAT_CHECK(keep[0] - hist_addr < (uint32_t)K3Mode<MODE>::kTableKB * 1024u && keep[1] - hist_addr < (uint32_t)K3Mode<MODE>::kTableKB * 1024u &&
                 keep[2] - hist_addr < (uint32_t)K3Mode<MODE>::kTableKB * 1024u && keep[3] - hist_addr < (uint32_t)K3Mode<MODE>::kTableKB * 1024u);
    k3_red(keep[0], k3_inc<0>(H, E));
    k3_red(keep[1], k3_inc<1>(H, E));
    k3_red(keep[2], k3_inc<2>(H, E));
    k3_red(keep[3], k3_inc<3>(H, E));
}
template <int MODE>
__device__ __forceinline__ void k3_take4(uint32_t hist_addr, uint32_t I4, uint32_t J4, uint32_t& sq, uint32_t& cnt) {
    constexpr int SH = K3Mode<MODE>::kAddrShift;
    uint32_t X, R, H, E;
    k3_pack4<MODE>(I4, J4, 0u, X, R, H, E);
    k3_take<MODE>(k3_addr<0, SH>(hist_addr, X, R), sq, cnt);
    k3_take<MODE>(k3_addr<1, SH>(hist_addr, X, R), sq, cnt);
    k3_take<MODE>(k3_addr<2, SH>(hist_addr, X, R), sq, cnt);
    k3_take<MODE>(k3_addr<3, SH>(hist_addr, X, R), sq, cnt);
}
// parity dump: the bins the table holds in this mode, as 32-bit counts at dump[i * 256 + j]
template <int MODE, int NT>
__device__ __forceinline__ void k3_dump_table(const uint32_t* hist, uint32_t* dump) {
    constexpr int words = K3Mode<MODE>::kTableKB * 256;
    for (int k = threadIdx.x; k < words; k += NT) {
        const uint32_t wv = hist[k], uk = (uint32_t)k;
        if (MODE == 0) {
            const uint32_t i = uk >> 6, jl = (uk & 63u) ^ ((i & 15u) << 2);
#pragma unroll
            for (int q = 0; q < 4; ++q) dump[i * 256u + jl + 64u * q] = (wv >> (8 * q)) & 0xffu;
        } else if (MODE <= 2) {
            const uint32_t i = 2u * (uk >> 7) + (MODE == 2 ? 1u : 0u), j0 = (uk & 127u) * 2u;
            reinterpret_cast<uint2*>(dump)[(i * 256u + j0) >> 1] = make_uint2(wv & 0xffffu, wv >> 16);
        } else if (MODE == 3) {
            const uint32_t i = uk >> 5, jl = (uk & 31u) ^ ((i & 7u) << 2);
#pragma unroll
            for (int q = 0; q < 8; ++q) dump[i * 256u + jl + 32u * q] = (wv >> (4 * q)) & 0xfu;
        } else if (MODE <= 5) {
            const uint32_t ih = uk >> 6, jl = (uk & 63u) ^ ((ih & 15u) << 2), i = 2u * ih + (MODE == 5 ? 1u : 0u);
#pragma unroll
            for (int q = 0; q < 4; ++q) dump[i * 256u + jl + 64u * q] = (wv >> (8 * q)) & 0xffu;
        } else {
            const uint32_t i = 4u * (uk >> 7) + (uint32_t)(MODE - 6), j0 = (uk & 127u) * 2u;
            reinterpret_cast<uint2*>(dump)[(i * 256u + j0) >> 1] = make_uint2(wv & 0xffffu, wv >> 16);
        }
    }
}

// ---- fallback passes: wider counters over a slice of the key space each --------------------------------------
// Only reached when a counter of the first attempt wrapped.  Adds the squared counts to S.slow[0] and the
// counts taken back to S.slow[2] (equal to the number of pairs unless a counter of these passes wrapped too).
template <bool MASKED, bool DUMP, int NT, int NG, int MODE>
__device__ __forceinline__ void k3_slow_pass(K3Smem& S, const uint32_t* hist, const K3Group& Gp, const K3Geom& G,
                                             uint32_t hist_addr, uint32_t* dump, uint32_t& sq, uint32_t& cnt) {
    const int tid = threadIdx.x;
    for (int item = tid; item < G.items; item += NT) {
        uint32_t I4[NG], J4[NG], pm, unused[4];
        if (!k3_item<MASKED, NG>(Gp, G, item, I4, J4, pm)) continue;
#pragma unroll
        for (int k = 0; k < NG; ++k)
            k3_build4<MODE>(hist_addr, I4[k], J4[k], (((pm >> (4 * k)) & 0xfu) * 0x00204081u) & 0x01010101u, unused);
    }
    __syncthreads();
    if (DUMP) {
        k3_dump_table<MODE, NT>(hist, dump);
        __syncthreads();
    }
    for (int item = tid; item < G.items; item += NT) {
        uint32_t I4[NG], J4[NG], pm;
        if (!k3_item<MASKED, NG>(Gp, G, item, I4, J4, pm)) continue;
#pragma unroll
        for (int k = 0; k < NG; ++k) k3_take4<MODE>(hist_addr, I4[k], J4[k], sq, cnt);
    }
    __syncthreads();
}
template <bool MASKED, bool DUMP, int NT, int NG, int FIRST, int NPASS>
__device__ __noinline__ void k3_slow_direction(K3Smem& S, const uint32_t* hist, const K3Group& Gp, const K3Geom& G,
                                               uint32_t hist_addr, uint32_t* dump) {
    uint32_t sq = 0u, cnt = 0u;
    k3_slow_pass<MASKED, DUMP, NT, NG, FIRST>(S, hist, Gp, G, hist_addr, dump, sq, cnt);
    k3_slow_pass<MASKED, DUMP, NT, NG, FIRST + 1>(S, hist, Gp, G, hist_addr, dump, sq, cnt);
    if (NPASS == 4) {
        k3_slow_pass<MASKED, DUMP, NT, NG, FIRST + 2>(S, hist, Gp, G, hist_addr, dump, sq, cnt);
        k3_slow_pass<MASKED, DUMP, NT, NG, FIRST + 3>(S, hist, Gp, G, hist_addr, dump, sq, cnt);
    }
    sq = __reduce_add_sync(0xffffffffu, sq);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) {
        if (sq) atomicAdd(&S.slow[0], sq);
        if (cnt) atomicAdd(&S.slow[2], cnt);
    }
    __syncthreads();
}

// ---- K3a, the front kernel: one WARP per tile, no CTA-wide barrier ------------------------------------------
// Quantises the tile (maximum from K1's column, or from a first pass when the basic block is off), packs
// the mask bits and their bounding box, works out the geometry of every direction, adds up the pair-stream
// sums of every direction into the output records, and leaves header + quantised pixels + mask bits in the
// scratch record of the tile for the bins kernel.
#ifndef IMFEAT_K3A_CTAS
#define IMFEAT_K3A_CTAS 8               // resident CTAs per SM the register budget is set for
#endif
template <bool MASKED>
__global__ void __launch_bounds__(32 * kK3aWarps, IMFEAT_K3A_CTAS)
k3a_front_kernel(const __grid_constant__ Params P, K3Cap max_pixels, uint32_t tile_base, uint32_t n_local_host,
                 unsigned char* __restrict__ scratch, int prefetch, K3Tier tier) {
    extern __shared__ __align__(16) unsigned char k3a_smem_raw[];
    __shared__ double homtab[256];
    constexpr int NG = MASKED ? 2 : 4;           // groups of 4 pairs per item
    constexpr int NP = 4 * NG;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_local = tier.wl_in ? min(*tier.n_in, n_local_host) : n_local_host;
    const size_t rec_bytes = k3_rec_bytes(max_pixels, MASKED);
    unsigned char* mine = k3a_smem_raw + (size_t)warp * k3a_warp_bytes(max_pixels, MASKED);
    K3Hdr& H = *reinterpret_cast<K3Hdr*>(mine);
    K3Group Gp;
    Gp.mbits = reinterpret_cast<uint32_t*>(mine + sizeof(K3Hdr));
    Gp.q8 = Gp.mbits + k3_mb_words(max_pixels.mb, MASKED);
#ifdef IMFEAT_CHECKS
    Gp.q8_words = k3_q8_words(max_pixels.q8); Gp.mb_words = k3_mb_words(max_pixels.mb, MASKED);
#endif
    uint32_t* recs = reinterpret_cast<uint32_t*>(mine + rec_bytes);      // [n_angles][kK3Rec]
    uint32_t* const rec_len = reinterpret_cast<uint32_t*>(scratch + (size_t)n_local_host * rec_bytes);
    auto tile_of = [&](long long tl) -> long long {          // global tile index of local tile tl
        return tier.wl_in ? (long long)tier.wl_in[tl] : (long long)tile_base + tl;
    };
    for (int k = threadIdx.x; k < 256; k += blockDim.x) homtab[k] = 1.0 / (1.0 + (double)(k * k));
    __syncthreads();
    const bool k1_max = P.col_basic >= 0;
    auto k1_max_of = [&](long long tl) -> double {            // K1 (earlier launch, same stream) left the maximum in the table
        if (!k1_max || tl >= (long long)n_local) return 0.0;
        const uint32_t tile = (uint32_t)tile_of(tl);
        const uint32_t row = tile / (uint32_t)P.c_out, slot = tile - row * (uint32_t)P.c_out;
        return __ldg(P.out + (long long)row * P.row_stride + P.col_basic + kNBasic * (int)slot + 10);
    };

    long long tnext = next_tile(tier.counter), tnext2 = next_tile(tier.counter);
    double vnext = k1_max_of(tnext);
    while (tnext < (long long)n_local) {
        const uint32_t tl = (uint32_t)tnext;
        const double vmaxd = vnext;
        tnext = tnext2;
        tnext2 = next_tile(tier.counter);                      // two tiles ahead: the next one is known now ...
        vnext = k1_max_of(tnext);
        if (prefetch && tnext < (long long)n_local) prefetch_tile_l2(P, tile_of(tnext), P.n_tiles);   // ... and on its way into L2
        const long long tile_id = tile_of(tl);
        const Tile T = resolve_tile(P, tile_id);
        const int tw = T.w, th = T.h, tn = T.n;
        const uint4* px4 = reinterpret_cast<const uint4*>(T.px);
        const uint2* mk2 = reinterpret_cast<const uint2*>(T.mk);
        const int nfull = tn >> 3, rem = tn & 7;
        uint8_t* mbytes = reinterpret_cast<uint8_t*>(Gp.mbits);
        uint32_t mul = 0u, sh = 24u;
        bool fast = false;
        if (k1_max) {
            const uint32_t vmax = (vmaxd == vmaxd) ? (uint32_t)vmaxd : 0u;   // NaN: empty mask, no pair exists anyway
            fast = k3_magic_fast(vmax, mul);
            if (!fast) k3_magic(vmax, mul, sh);
        } else {
            // no basic block in this call: the maximum (over the mask) from a first pass of our own
            uint32_t mx2 = 0u;
            for (int idx = lane; idx < nfull; idx += 32) {
                uint4 v = ld_reuse(px4 + idx);
                if (MASKED) {
                    const uint2 m = __ldg(mk2 + idx);
                    const uint32_t c0 = __vcmpne4(m.x, 0u), c1 = __vcmpne4(m.y, 0u);
                    v.x &= __byte_perm(c0, 0u, 0x1100); v.y &= __byte_perm(c0, 0u, 0x3322);
                    v.z &= __byte_perm(c1, 0u, 0x1100); v.w &= __byte_perm(c1, 0u, 0x3322);
                }
                mx2 = __vmaxu2(mx2, __vmaxu2(__vmaxu2(v.x, v.y), __vmaxu2(v.z, v.w)));
            }
            if (lane < rem && (!MASKED || T.mk[nfull * 8 + lane] != 0)) mx2 = __vmaxu2(mx2, (uint32_t)T.px[nfull * 8 + lane]);
            const uint32_t vmax = __reduce_max_sync(0xffffffffu, max(mx2 & 0xffffu, mx2 >> 16));
            fast = k3_magic_fast(vmax, mul);
            if (!fast) k3_magic(vmax, mul, sh);
        }

        // ---- stage the tile: mask bits + bounding box first, then the 8-bit quantisation of the rows the box spans ----
        int brmin = 1 << 30, brmax = -1, bcmin = 1 << 30, bcmax = -1;
        const float rtw = __frcp_rn((float)tw);
        auto mask_chunk = [&](int idx, const uint2& m) {
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
        };
        if (MASKED) {
            constexpr int kUm = 8;                         // mask loads in flight per lane
            int idx = lane;
            for (; idx + 32 * (kUm - 1) < nfull; idx += 32 * kUm) {
                uint2 m[kUm];
#pragma unroll
                for (int u = 0; u < kUm; ++u) m[u] = __ldg(mk2 + idx + 32 * u);
#pragma unroll
                for (int u = 0; u < kUm; ++u) mask_chunk(idx + 32 * u, m[u]);
            }
            for (; idx < nfull; idx += 32) mask_chunk(idx, __ldg(mk2 + idx));
            if (lane == 0 && rem) {                        // tail pixels (< 8): one lane, in order
                uint32_t bits = 0u;
                for (int k = 0; k < rem; ++k) {
                    const int i = nfull * 8 + k;
                    if (T.mk[i] != 0) {
                        bits |= 1u << k;
                        const int ra = i / tw, ca = i - ra * tw;
                        brmin = min(brmin, ra); brmax = max(brmax, ra); bcmin = min(bcmin, ca); bcmax = max(bcmax, ca);
                    }
                }
                mbytes[nfull] = (uint8_t)bits;
            }
            brmin = __reduce_min_sync(0xffffffffu, brmin); brmax = __reduce_max_sync(0xffffffffu, brmax);
            bcmin = __reduce_min_sync(0xffffffffu, bcmin); bcmax = __reduce_max_sync(0xffffffffu, bcmax);
            if (brmax < 0) { brmin = 0; bcmin = 0; bcmax = -1; }           // empty mask: nothing to stage, no pair
        } else {
            brmin = 0; brmax = th - 1; bcmin = 0; bcmax = tw - 1;
        }
        const int pitch = k3_pitch<MASKED>(tw), cpr = tw >> 3;         // padded pitch: 8-pixel chunks per row
        // the staged rows: pixel (r, c) lies at byte r * pitch + c - qbias of the quantised block
        const int p_lo = brmin * tw, p_hi = (brmax + 1) * tw;
        const int i0 = p_lo >> 3, i1 = min(nfull, (p_hi + 7) >> 3);
        const int qbias = pitch == tw ? (p_lo & ~7) : brmin * pitch;
        const uint32_t q8_used = (uint32_t)((pitch == tw ? p_hi : (brmax + 1) * pitch) - qbias);
        if (tier.wl_out && q8_used > (uint32_t)max_pixels.q8) {
            // the rows of this mask's bounding box do not fit this tier: leave the tile to the next one
            if (lane == 0) {
                const uint32_t wpos = atomicAdd(tier.n_out, 1u);
                IMFEAT_CHECK(wpos < n_local_host);
                tier.wl_out[wpos] = (uint32_t)tile_id;
                rec_len[tl] = kK3Skip;
            }
            __syncwarp();
            continue;
        }
        const float rcpr = __frcp_rn((float)max(cpr, 1));
        auto quant_chunk = [&](int idx, const uint4& v) {              // out-of-mask pixels may exceed the maximum:
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};               // their bytes are never part of a pair
            uint32_t q[4];
            if (fast) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    q[k] = __byte_perm(__umulhi(w4[k] & 0xffffu, mul), __umulhi(w4[k] >> 16, mul), 0x0040);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    q[k] = __byte_perm(k3_quant(w4[k] & 0xffffu, mul, sh), k3_quant(w4[k] >> 16, mul, sh), 0x0040);
            }
            const uint32_t lo = __byte_perm(q[0], q[1], 0x5410), hi = __byte_perm(q[2], q[3], 0x5410);
            if (pitch == tw) {
                IMFEAT_CHECK(2 * idx - (qbias >> 2) >= 0 && 2 * idx - (qbias >> 2) + 1 < k3_q8_words(max_pixels.q8));
                *reinterpret_cast<uint2*>(Gp.q8 + 2 * idx - (qbias >> 2)) = make_uint2(lo, hi);
            } else {                                       // the chunk lies inside one row (tw % 8 == 0)
                const int ra = (int)(((float)idx + 0.5f) * rcpr);
                IMFEAT_CHECK(ra >= brmin && (((ra - brmin) * pitch) >> 2) + 2 * (idx - ra * cpr) + 1 < k3_q8_words(max_pixels.q8));
                uint32_t* dst = Gp.q8 + (((ra - brmin) * pitch) >> 2) + 2 * (idx - ra * cpr);
                dst[0] = lo; dst[1] = hi;
            }
        };
        {
            constexpr int kU = 4;                          // loads in flight per lane
            int idx = i0 + lane;
            for (; idx + 32 * (kU - 1) < i1; idx += 32 * kU) {
                uint4 v[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) v[u] = k1_max ? ld_stream(px4 + idx + 32 * u) : ld_reuse(px4 + idx + 32 * u);
#pragma unroll
                for (int u = 0; u < kU; ++u) quant_chunk(idx + 32 * u, v[u]);
            }
            for (; idx < i1; idx += 32) quant_chunk(idx, ld_reuse(px4 + idx));
        }
        if (lane < rem && p_hi > nfull * 8) {
            const int i = nfull * 8 + lane;
            const uint32_t qv = fast ? __umulhi((uint32_t)T.px[i], mul) : k3_quant(T.px[i], mul, sh);
            reinterpret_cast<uint8_t*>(Gp.q8)[i - qbias] = (uint8_t)qv;
        }
        // geometry of the directions: lane a works out direction a
        if (lane < P.n_angles) H.geom[lane] = k3_geom<MASKED, NG>(tw, P.dr[lane], P.dc[lane], brmin, brmax, bcmin, bcmax, qbias);
        const uint32_t len = ((uint32_t)sizeof(K3Hdr) + 4u * (uint32_t)k3_mb_words(max_pixels.mb, MASKED) + q8_used + 15u) & ~15u;
        IMFEAT_CHECK(len <= (uint32_t)rec_bytes && (int)q8_used <= max_pixels.q8 + 32);
        if (lane == 0) {
            H.rec = reinterpret_cast<uint32_t*>(T.out_row + P.col_glcm + T.slot * P.n_angles * kNGlcm);
            H.tile = (uint32_t)tile_id; H.len = len;
            rec_len[tl] = len;
        }
        __syncwarp();

        // ---- pair-stream sums of every direction ----
#pragma unroll 1
        for (int a = 0; a < P.n_angles; ++a) {
            const K3Geom G = H.geom[a];
            K3Acc A = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0.0, 0.0};
            uint32_t np = 0u;
            for (int item = lane; item < G.items; item += 32) {
                uint32_t I4[NG], J4[NG], pm;
                if (k3_item<MASKED, NG>(Gp, G, item, I4, J4, pm)) {
#pragma unroll
                    for (int k = 0; k < NG; ++k)
                        k3_sums(homtab, I4[k], J4[k], ((((pm >> (4 * k)) & 0xfu) * 0x00204081u) & 0x01010101u) * 0xffu, A);
                    A.m += __popc(pm);
                    np += NP;
                }
            }
            const uint32_t r0 = __reduce_add_sync(0xffffffffu, A.si);
            const uint32_t r1 = __reduce_add_sync(0xffffffffu, A.sj);
            const uint32_t r2 = __reduce_add_sync(0xffffffffu, A.sii);
            const uint32_t r3 = __reduce_add_sync(0xffffffffu, A.sjj);
            const uint32_t r4 = __reduce_add_sync(0xffffffffu, A.sij);
            const uint32_t r5 = __reduce_add_sync(0xffffffffu, A.sd);
            const uint32_t r7 = __reduce_add_sync(0xffffffffu, A.m);
            const uint32_t r8 = __reduce_add_sync(0xffffffffu, np);
            // per-lane double sums (fixed order) are rounded to 2^-40 fixed point: the sum over lanes is integer
            const unsigned long long hf = warp_sum_redux((unsigned long long)__double2ll_rn((A.hom0 + A.hom1) * 1099511627776.0));
            if (lane == 0) {
                uint32_t* r = recs + a * kK3Rec;
                r[kR_si] = r0; r[kR_sj] = r1; r[kR_sii] = r2; r[kR_sjj] = r3; r[kR_sij] = r4; r[kR_sd] = r5;
                r[kR_sq] = 0u; r[kR_m] = r7; r[kR_homlo] = (uint32_t)hf; r[kR_homhi] = (uint32_t)(hf >> 32);
                r[kR_np] = r8; r[kR_cnt] = 0u;
                H.geom[a].pad[0] = (int)r7;                // the bins kernel checks its counts against this
            }
        }
        __syncwarp();
        // the records of all directions are contiguous in the output row; the scratch record is one block
        for (int k = lane; k < P.n_angles * kK3Rec; k += 32) H.rec[k] = recs[k];
        {
            const uint4* src = reinterpret_cast<const uint4*>(mine);
            uint4* dst = reinterpret_cast<uint4*>(scratch + (size_t)tl * rec_bytes);
            for (int k = lane; k < (int)(len >> 4); k += 32) dst[k] = src[k];
        }
        __syncwarp();                                      // the buffers are rewritten by the next tile
    }
}

// ---- K3b, the bins kernel ------------------------------------------------------------------------------------
// TB = 64: 8-bit counters first (MODE 0), 16-bit halves if one wrapped.  TB = 32: 4-bit counters first
// (MODE 3), 8-bit halves if one wrapped, 16-bit quarters if one of those wrapped too; a CTA that has just seen
// a wrap starts the next directions with the 8-bit halves at once.
template <bool MASKED, bool DUMP, int NT, int TB>
__global__ void __launch_bounds__(NT, TB == 64 ? 3 : 5)
k3_glcm_kernel(const __grid_constant__ Params P, K3Cap max_pixels, uint32_t n_local_host, const unsigned char* __restrict__ scratch,
               K3Tier tier) {
    extern __shared__ __align__(16) unsigned char k3_smem_raw[];
    uint32_t* const hist = reinterpret_cast<uint32_t*>(k3_smem_raw);
    K3Smem& S = *reinterpret_cast<K3Smem*>(k3_smem_raw + TB * 1024);
    constexpr int M0 = TB == 64 ? 0 : 3;         // the mode of the first attempt
    constexpr int NG = MASKED ? 2 : 4;           // groups of 4 pairs per item
    constexpr int NP = 4 * NG;
    constexpr int KC = 4096 / (NT * NP);         // items per thread whose word addresses stay in registers (a 64x64 tile in full)
    constexpr int NW = NT / 32;
    static_assert(NW <= kK3MaxWarps && KC >= 1, "CTA size");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n_local = tier.wl_in ? min(*tier.n_in, n_local_host) : n_local_host;
    if (blockIdx.x >= n_local) return;
    const uint32_t rec_bytes = (uint32_t)k3_rec_bytes(max_pixels, MASKED);
    unsigned char* recbuf0 = k3_smem_raw + TB * 1024 + sizeof(K3Smem);
    const uint32_t hist_addr = smem_addr(hist), bar0 = smem_addr(&S.mbar[0]), rec_addr0 = smem_addr(recbuf0);
    const uint32_t* const rec_len = reinterpret_cast<const uint32_t*>(scratch + (size_t)n_local_host * rec_bytes);

    for (int k = tid; k < TB * 256; k += NT) hist[k] = 0u;
    if (tid == 0) {
        // tiles: the first one is the CTA's own index, the others come from the per-launch counter (so the
        // tail does not depend on how many CTAs are resident at once), drawn two tiles ahead
        S.tq[0] = blockIdx.x;
        S.tq[1] = gridDim.x + atomicAdd(tier.counter, 1u);
        S.tlen[0] = blockIdx.x < n_local ? rec_len[blockIdx.x] : 0u;
        S.tlen[1] = S.tq[1] < n_local ? rec_len[S.tq[1]] : 0u;
        S.slow[0] = 0u; S.slow[1] = 0u; S.slow[2] = 0u;
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // the record of a tile (header, quantised pixels, mask bits) arrives with one bulk copy (1-D TMA) in one of two
    // buffers: the record of tile j + 1 travels while tile j is worked off
    auto fetch = [&](uint32_t tl, uint32_t b, uint32_t len) { // thread 0
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // earlier generic reads of the buffer vs the async write
        IMFEAT_CHECK(len >= kK3Skip && len <= rec_bytes && (len & 15u) == 0u && tl < n_local_host);
        mbar_expect_tx(bar0 + 8 * b, len);
        bulk_g2s(rec_addr0 + b * rec_bytes, scratch + (size_t)tl * rec_bytes, len, bar0 + 8 * b);
    };
    if (tid == 0 && blockIdx.x < n_local) fetch(blockIdx.x, 0u, S.tlen[0]);
    uint32_t dcount = 0u;                        // directions done by this CTA: its parity picks the partial-sum bank
    int prefer_wide = 0;                         // TB = 32: directions left that skip the 4-bit attempt

    for (uint32_t j = 0;; ++j) {
        const uint32_t tl = S.tq[j & 3u];
        if (tl >= n_local) break;
        const uint32_t b = j & 1u;
        uint32_t t_draw = 0u, len_draw = 0u;
        if (tid == 0) {
            t_draw = gridDim.x + atomicAdd(tier.counter, 1u);
            const uint32_t t_next = S.tq[(j + 1) & 3u];
            if (t_next < n_local) fetch(t_next, b ^ 1u, S.tlen[(j + 1) & 3u]);     // that buffer's tile (j - 1) was finished before the last barrier
            if (t_draw < n_local) len_draw = rec_len[t_draw];
        }
        const K3Hdr& H = *reinterpret_cast<const K3Hdr*>(recbuf0 + b * rec_bytes);
        K3Group Gp;
        Gp.mbits = reinterpret_cast<uint32_t*>(recbuf0 + b * rec_bytes + sizeof(K3Hdr));
        Gp.q8 = Gp.mbits + k3_mb_words(max_pixels.mb, MASKED);
#ifdef IMFEAT_CHECKS
        Gp.q8_words = k3_q8_words(max_pixels.q8); Gp.mb_words = k3_mb_words(max_pixels.mb, MASKED);
#endif
        mbar_wait(bar0 + 8 * b, (j >> 1) & 1u);            // ---- record landed ----
        if (S.tlen[j & 3u] == kK3Skip) {                   // a tile left to the second tier: nothing to do here
            if (tid == 0) { S.tq[(j + 2) & 3u] = t_draw; S.tlen[(j + 2) & 3u] = len_draw; }
            __syncthreads();
            continue;
        }

#pragma unroll 1
        for (int a = 0; a < P.n_angles; ++a, ++dcount) {
            const K3Geom G = H.geom[a];
            const int par = (int)(dcount & 1u);
            const uint32_t M = (uint32_t)G.pad[0];         // pairs of this direction (front kernel)
            uint32_t* dump = nullptr;
            if (DUMP) dump = P.counts + ((long long)H.tile * P.n_angles + a) * 65536ll;
            uint32_t sqt = 0u, cntt = 0u;
            const bool first = !(TB == 32 && prefer_wide > 0);
            if (first) {
                uint32_t adr[KC][NP];
                uint32_t valid = 0u;
                // ---- build: every existing pair adds 1 to its bin ----
#pragma unroll
                for (int i = 0; i < KC; ++i) {
                    const int item = tid + i * NT;
                    uint32_t I4[NG], J4[NG], pm;
                    if (item < G.items && k3_item<MASKED, NG>(Gp, G, item, I4, J4, pm)) {
#pragma unroll
                        for (int k = 0; k < NG; ++k) {
                            uint32_t k4[4];
                            k3_build4<M0>(hist_addr, I4[k], J4[k], (((pm >> (4 * k)) & 0xfu) * 0x00204081u) & 0x01010101u, k4);
                            adr[i][4 * k] = k4[0]; adr[i][4 * k + 1] = k4[1]; adr[i][4 * k + 2] = k4[2]; adr[i][4 * k + 3] = k4[3];
                        }
                        valid |= 1u << i;
                    }
                }
                for (int item = tid + KC * NT; item < G.items; item += NT) {
                    uint32_t I4[NG], J4[NG], pm, unused[4];
                    if (k3_item<MASKED, NG>(Gp, G, item, I4, J4, pm)) {
#pragma unroll
                        for (int k = 0; k < NG; ++k)
                            k3_build4<M0>(hist_addr, I4[k], J4[k], (((pm >> (4 * k)) & 0xfu) * 0x00204081u) & 0x01010101u, unused);
                    }
                }
                __syncthreads();                           // ---- bins of this direction complete ----
                if (DUMP) {
                    // parity dump of the raw bins: this table if no counter wrapped (checked densely here),
                    // else from the passes of the fallback below
                    uint32_t dsum = 0u;
                    for (int k = tid; k < TB * 256; k += NT) {
                        const uint32_t wv = hist[k];
                        dsum = TB == 64 ? __dp4a(wv, 0x01010101u, dsum)
                                        : __dp4a((wv & 0x0f0f0f0fu) + ((wv >> 4) & 0x0f0f0f0fu), 0x01010101u, dsum);
                    }
                    dsum = __reduce_add_sync(0xffffffffu, dsum);
                    if (lane == 0 && dsum) atomicAdd(&S.slow[1], dsum);
                    __syncthreads();
                    if (S.slow[1] == M) k3_dump_table<M0, NT>(hist, dump);
                    __syncthreads();
                    if (tid == 0) S.slow[1] = 0u;
                }
                // ---- clear: take the counts back, summing their squares ----
                uint32_t sq = 0u, cnt = 0u;
#pragma unroll
                for (int i = 0; i < KC; ++i)
                    if (valid & (1u << i)) {
#pragma unroll
                        for (int k = 0; k < NP; ++k) k3_take<M0>(adr[i][k], sq, cnt);
                    }
                for (int item = tid + KC * NT; item < G.items; item += NT) {
                    uint32_t I4[NG], J4[NG], pm;
                    if (k3_item<MASKED, NG>(Gp, G, item, I4, J4, pm)) {
#pragma unroll
                        for (int k = 0; k < NG; ++k) k3_take4<M0>(hist_addr, I4[k], J4[k], sq, cnt);
                    }
                }
                sq = __reduce_add_sync(0xffffffffu, sq);
                cnt = __reduce_add_sync(0xffffffffu, cnt);
                if (lane == 0) { S.part[par][warp][0] = sq; S.part[par][warp][1] = cnt; }
                if (tid == 0 && a + 1 == P.n_angles) { S.tq[(j + 2) & 3u] = t_draw; S.tlen[(j + 2) & 3u] = len_draw; }
                __syncthreads();                           // ---- table clean, sums complete ----
#pragma unroll
                for (int w = 0; w < NW; ++w) { sqt += S.part[par][w][0]; cntt += S.part[par][w][1]; }
            } else {
                --prefer_wide;
                if (tid == 0 && a + 1 == P.n_angles) { S.tq[(j + 2) & 3u] = t_draw; S.tlen[(j + 2) & 3u] = len_draw; }
                cntt = M + 1u;                             // straight to the wider counters
            }
            if (cntt != M) {
                // a counter wrapped (a carry changes the sum of the counts): exact passes with wider counters
                if (TB == 64) {
                    k3_slow_direction<MASKED, DUMP, NT, NG, 1, 2>(S, hist, Gp, G, hist_addr, dump);
                } else {
                    if (first) prefer_wide = 16;
                    if (!DUMP) k3_slow_direction<MASKED, false, NT, NG, 4, 2>(S, hist, Gp, G, hist_addr, dump);
                    const bool again = DUMP || S.slow[2] != M;
                    __syncthreads();
                    if (again) {
                        if (tid == 0) { S.slow[0] = 0u; S.slow[2] = 0u; }
                        __syncthreads();
                        k3_slow_direction<MASKED, DUMP, NT, NG, 6, 4>(S, hist, Gp, G, hist_addr, dump);
                    }
                }
                sqt = S.slow[0];
                __syncthreads();
                if (tid == 0) { S.slow[0] = 0u; S.slow[2] = 0u; }
            }
            if (tid == 0) H.rec[a * kK3Rec + kR_sq] = sqt;
        }
    }
}

// One direction's six properties from the exact integer sums the table kernel left in the output
// columns (in place: a thread reads its own 48-byte record, then overwrites it).
__global__ void __launch_bounds__(256) k3_finalize_kernel(const __grid_constant__ Params P) {
    const long long total = P.n_tiles * P.n_angles;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const uint32_t tile = (uint32_t)(g / P.n_angles);
        const int a = (int)(g - (long long)tile * P.n_angles);
        const uint32_t row = tile / (uint32_t)P.c_out, slot = tile - row * (uint32_t)P.c_out;
        double* o = P.out + (long long)row * P.row_stride + P.col_glcm + ((int)slot * P.n_angles + a) * kNGlcm;
        uint32_t r[kK3Rec];
#pragma unroll
        for (int k = 0; k < kK3Rec / 2; ++k) {
            const uint2 v = reinterpret_cast<const uint2*>(o)[k];
            r[2 * k] = v.x; r[2 * k + 1] = v.y;
        }
        const long long M = r[kR_m];
        if (M == 0) {
            o[0] = 0.0; o[1] = 0.0; o[2] = 0.0; o[3] = 0.0; o[4] = 0.0; o[5] = 1.0;
            if (P.status) atomicOr(P.status + row, kStNoPairs);
            continue;
        }
        // the walked pairs that do not exist added exactly 1.0 each to the homogeneity sum
        const unsigned long long hom_true = ((unsigned long long)r[kR_homhi] << 32 | r[kR_homlo]) -
                                            ((unsigned long long)((long long)r[kR_np] - M) << 40);
        const double Md = (double)M;
        const long long Si = r[kR_si], Sj = r[kR_sj], Sii = r[kR_sii], Sjj = r[kR_sjj], Sij = r[kR_sij];
        const long long vi = M * Sii - Si * Si, vj = M * Sjj - Sj * Sj, cov = M * Sij - Si * Sj;
        const double asmv = (double)r[kR_sq] / (Md * Md);
        o[0] = (double)(Sii + Sjj - 2 * Sij) / Md;
        o[1] = (double)r[kR_sd] / Md;
        o[2] = ((double)hom_true * 9.094947017729282e-13) / Md;
        o[3] = asmv;
        o[4] = sqrt(asmv);
        o[5] = (vi == 0 || vj == 0) ? 1.0 : (double)cov / (sqrt((double)vi) * sqrt((double)vj));
    }
}

}  // namespace imfeat
