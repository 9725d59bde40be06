// k3_glcm.cuh -- K3: gray-level quantisation + GLCM + Haralick properties.
//
// Replaces, per channel:
//   (x / x.max()) * 255 -> uint8                      NB:293-295
//   greycomatrix(q, [5], [0], levels=256)              NB:298   (not symmetric, not normed)
//   greycoprops x6 (contrast .. correlation)           NB:301-306
//
// * quantiser: floor(255*x/max) with an exact multiply-shift reciprocal (bit-identical to the
//   notebook's float64 expression for every uint16 pair; tests/test_oracle_cpu.py).
// * the 256x256 bins live in shared memory as 16-bit counters (two per 32-bit word, 128 KB),
//   built with shared-memory atomics on the pair stream, dumped on request (parity), and only
//   ever cleared sparsely by re-walking the pairs.  Two 512-thread groups work on two tiles at a
//   time and take turns on the table (common.cuh, "ping-pong"); everything that does not need
//   the table (loads, max, quantisation, pair-stream sums, reductions, epilogue) overlaps the
//   other group's table phase.
// * contrast / dissimilarity / correlation come from exact integer sums over the pair stream,
//   four pairs per SIMD video instruction (dp4a, vabsdiff4); ASM = sum_bins c^2 is accumulated
//   from the atomics' return values (c^2 = sum_{k<c} (2k+1) = 2*sum(old) + c), so there is no
//   pass over the bins and no read-back pass.
#pragma once
#include "common.cuh"

namespace imfeat {


constexpr int kK3Cache = 4;     // pair groups (4 pairs each) per thread cached in registers

// Shared-memory layout (dynamic): [hist 128 KB][homtab 2 KB][tokens][per-group: K3GroupHdr, q8, mbits]
struct K3GroupHdr {
    // per-tile scratch exists twice and alternates: a tile's epilogue runs while the next tile is
    // already under way (after that tile's staging barrier), so no barrier is spent on it
    unsigned long long whom[2][kMaxAngles][16];   // per-warp sums of 1/(1+d^2) in 2^-40 fixed point
    uint32_t acc[2][kMaxAngles][8];               // si sj sii sjj sij sd sold m, per direction
    int box[2][4];                                // mask bounding box: rmin, rmax, cmin, cmax (masked variant)
    uint32_t wmax[32];
};
struct K3Smem {
    uint32_t hist[32768];
    uint32_t dummy[34];      // at hist + 0x20000: word 0 takes the non-existent pairs of the unmasked path,
                             // words 2..33 (one per lane) those of the masked path
    double homtab[256];
    unsigned long long tokens[8];
};
struct K3Group {                               // pointers into the dynamic region of this group
    K3GroupHdr* hdr;
    uint32_t* q8;                              // quantised pixels (bytes) + slack for unaligned reads
    uint32_t* mbits;                           // one bit per pixel: inside the mask (masked variant)
};
__host__ __device__ inline size_t k3_group_bytes(int max_pixels) {
    const size_t q8w = (size_t)max_pixels / 4 + 4, mbw = (size_t)max_pixels / 32 + 2;
    return sizeof(K3GroupHdr) + 4 * ((q8w + 1) & ~(size_t)1) + 4 * ((mbw + 1) & ~(size_t)1);
}
__host__ __device__ inline size_t k3_smem_bytes(int max_pixels, int ng) {
    return sizeof(K3Smem) + (size_t)ng * k3_group_bytes(max_pixels);
}

struct K3Acc {
    uint32_t si, sj, sii, sjj, sij, sd, sasm, m;
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
__device__ __forceinline__ uint32_t k3_quant(uint32_t x, uint32_t mul, uint32_t sh) {
    return (uint32_t)(((unsigned long long)(x * 255u) * mul) >> sh);
}

// four consecutive bytes starting at byte offset off (any alignment)
__device__ __forceinline__ uint32_t k3_load4(const uint32_t* b, int off) {
    const int w = off >> 2;
    return __funnelshift_r(b[w], b[w + 1], (off & 3) << 3);
}
// four consecutive mask bits starting at bit offset off, expanded to 0xff / 0x00 bytes
__device__ __forceinline__ uint32_t k3_mask4(const uint32_t* b, int off) {
    const int w = off >> 5;
    const uint32_t bits = __funnelshift_r(b[w], b[w + 1], off & 31) & 0xfu;
    return ((bits * 0x00204081u) & 0x01010101u) * 0xffu;
}

struct K3Geom {
    int nrows, r0, c0, c1, gpr, lg, items, w, doff;
};
// Pairs (r, c) -> (r + dr, c + dc) with both pixels inside the box rows [br0, br1], columns
// [bc0, bc1] (the whole tile, or the bounding box of the mask: pairs outside it cannot exist).
__device__ __forceinline__ K3Geom k3_geom(int w, int dr, int dc, int br0, int br1, int bc0, int bc1) {
    K3Geom G;
    G.r0 = br0;
    G.nrows = br1 - dr - br0 + 1;                  // dr >= 0 for all supported directions
    G.c0 = bc0 + (dc < 0 ? -dc : 0);
    G.c1 = bc1 + 1 - (dc > 0 ? dc : 0);
    G.w = w;
    G.doff = dr * w + dc;
    if (G.nrows <= 0 || G.c1 <= G.c0) { G.items = 0; G.gpr = 0; G.lg = 0; G.nrows = 0; return G; }
    G.gpr = (G.c1 - G.c0 + 3) >> 2;                // groups of 4 pairs per row
    G.lg = G.gpr <= 1 ? 0 : 32 - __clz(G.gpr - 1); // ceil(log2 gpr): row index = item >> lg
    G.items = G.nrows << G.lg;
    return G;
}

// Load one item (4 horizontally consecutive pairs): quantised bytes of both pixels and the
// byte mask of the pairs that exist (inside the image, and inside the mask when MASKED).
template <bool MASKED>
__device__ __forceinline__ bool k3_item(const K3Group& Gp, const K3Geom& G, int item, uint32_t& I4,
                                        uint32_t& J4, uint32_t& vm) {
    const int r = G.r0 + (item >> G.lg), cg = item & ((1 << G.lg) - 1);
    if (cg >= G.gpr) { vm = 0u; I4 = J4 = 0u; return false; }
    const int c = G.c0 + 4 * cg;
    const int valid = min(4, G.c1 - c);
    const int oi = r * G.w + c, oj = oi + G.doff;
    vm = valid == 4 ? 0xffffffffu : ((1u << (8 * valid)) - 1u);
    if (MASKED) vm &= k3_mask4(Gp.mbits, oi) & k3_mask4(Gp.mbits, oj);
    I4 = k3_load4(Gp.q8, oi) & vm;
    J4 = k3_load4(Gp.q8, oj) & vm;
    return vm != 0u;
}

// Pair-stream sums of one item.  Unmasked tiles run branch-free: the bytes of non-existent pairs
// (row tails) are zero, so they add nothing to the integer sums and exactly homtab[0] = 1.0 to the
// homogeneity sum, which the epilogue subtracts again.
template <bool MASKED>
__device__ __forceinline__ void k3_sums(const K3Smem& S, uint32_t I4, uint32_t J4, uint32_t vm, K3Acc& A) {
    A.si = __dp4a(I4, 0x01010101u, A.si);
    A.sj = __dp4a(J4, 0x01010101u, A.sj);
    A.sii = __dp4a(I4, I4, A.sii);
    A.sjj = __dp4a(J4, J4, A.sjj);
    A.sij = __dp4a(I4, J4, A.sij);
    A.sd += __vsadu4(I4, J4);
    A.m += __popc(vm) >> 3;
    const uint32_t D4 = __vabsdiffu4(I4, J4);
    A.hom += S.homtab[D4 & 0xffu];
    A.hom += S.homtab[(D4 >> 8) & 0xffu];
    A.hom += S.homtab[(D4 >> 16) & 0xffu];
    A.hom += S.homtab[D4 >> 24];
    // masked tiles: take the 1.0 of every missing pair out again right away (exact)
    if (MASKED) A.hom -= (double)(4 - (__popc(vm) >> 3));
}

// PHASE 0: bins += 1, accumulating the returned old counts (sum_bins c^2 = 2*sum(old) + M);
// PHASE 2: sparse clear.  Keys (i << 8 | j) are assembled two at a time with PRMT.
// Both variants run branch-free.  Unmasked: a non-existent pair (row tail) is redirected to one
// dummy bin behind the table; D such pairs return the old values 0..D-1 in some order, so the
// epilogue subtracts D(D-1)/2.  Masked (many missing pairs): each lane has its own dummy word and
// the returned count of a missing pair is dropped with a select.
template <int PHASE, bool MASKED>
__device__ __forceinline__ void k3_bin1(K3Smem& S, uint32_t key, bool exists, uint32_t& sold) {
    uint32_t off = (key << 1) & 0x1fffcu;
    off = exists ? off : (MASKED ? 0x20008u + 4u * (threadIdx.x & 31) : 0x20000u);
    uint32_t* word = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(S.hist) + off);
    if (PHASE == 0) {
        const uint32_t sh = (key & 1u) << 4;       // key == 0 for a non-existent pair
        const uint32_t old = (atomicAdd(word, 1u << sh) >> sh) & 0xffffu;
        sold += (MASKED && !exists) ? 0u : old;    // unmasked: corrected by D(D-1)/2 in the epilogue
    } else {
        *word = 0u;
    }
}
template <int PHASE, bool MASKED>
__device__ __forceinline__ void k3_bins(K3Smem& S, uint32_t I4, uint32_t J4, uint32_t vm, uint32_t& sold) {
    const uint32_t K01 = __byte_perm(J4, I4, 0x5140);   // [j0, i0, j1, i1]
    const uint32_t K23 = __byte_perm(J4, I4, 0x7362);   // [j2, i2, j3, i3]
    k3_bin1<PHASE, MASKED>(S, K01 & 0xffffu, (vm & 0x00000001u) != 0u, sold);
    k3_bin1<PHASE, MASKED>(S, K01 >> 16, (vm & 0x00000100u) != 0u, sold);
    k3_bin1<PHASE, MASKED>(S, K23 & 0xffffu, (vm & 0x00010000u) != 0u, sold);
    k3_bin1<PHASE, MASKED>(S, K23 >> 16, (vm & 0x01000000u) != 0u, sold);
}

// One direction's six properties from the exact integer sums (s: si sj sii sjj sij sd sold m).
template <bool MASKED>
__device__ __forceinline__ void k3_epilogue(const Params& P, double* out_row, uint32_t* status, int slot,
                                            int h, int w, int a, const uint32_t* s, unsigned long long hom_sum) {
    const long long M = (long long)s[7];
    double* o = out_row + P.col_glcm + (slot * P.n_angles + a) * kNGlcm;
    if (M == 0) {
        o[0] = 0.0; o[1] = 0.0; o[2] = 0.0; o[3] = 0.0; o[4] = 0.0; o[5] = 1.0;
        if (status) atomicOr(status, kStNoPairs);
        return;
    }
    // non-existent pairs that went through the branch-free path (unmasked only)
    const K3Geom Ge = k3_geom(w, P.dr[a], P.dc[a], 0, h - 1, 0, w - 1);
    const long long D = MASKED ? 0ll : 4ll * Ge.nrows * Ge.gpr - M;
    const unsigned long long sold_true = (unsigned long long)s[6] - (unsigned long long)(D * (D - 1) / 2);
    const unsigned long long hom_true = hom_sum - ((unsigned long long)D << 40);
    const double Md = (double)M;
    const long long con = (long long)s[2] + (long long)s[3] - 2ll * (long long)s[4];
    const long long vi = M * (long long)s[2] - (long long)s[0] * (long long)s[0];
    const long long vj = M * (long long)s[3] - (long long)s[1] * (long long)s[1];
    const long long cov = M * (long long)s[4] - (long long)s[0] * (long long)s[1];
    const double asmv = (double)(2ull * sold_true + (unsigned long long)M) / (Md * Md);
    o[0] = (double)con / Md;
    o[1] = (double)s[5] / Md;
    o[2] = ((double)hom_true * 9.094947017729282e-13) / Md;
    o[3] = asmv;
    o[4] = sqrt(asmv);
    o[5] = (vi == 0 || vj == 0) ? 1.0 : (double)cov / (sqrt((double)vi) * sqrt((double)vj));
}

template <bool MASKED, bool DUMP>
__global__ void __launch_bounds__(1024, 1) k3_glcm_kernel(const __grid_constant__ Params P, int ng, int max_pixels) {
    extern __shared__ __align__(16) unsigned char k3_smem_raw[];
    K3Smem& S = *reinterpret_cast<K3Smem*>(k3_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    Ring R;
    ring_init(R, S.tokens, ng);
    const int g = R.g, gt = R.gt, gw = R.gw, gthreads = R.gthreads;
    K3Group Gp;
    {
        unsigned char* base = k3_smem_raw + sizeof(K3Smem) + (size_t)g * k3_group_bytes(max_pixels);
        const size_t q8w = ((size_t)max_pixels / 4 + 4 + 1) & ~(size_t)1;
        Gp.hdr = reinterpret_cast<K3GroupHdr*>(base);
        Gp.q8 = reinterpret_cast<uint32_t*>(base + sizeof(K3GroupHdr));
        Gp.mbits = Gp.q8 + q8w;
    }
    K3GroupHdr& H = *Gp.hdr;

    for (int k = tid; k < 32768; k += blockDim.x) S.hist[k] = 0u;
    if (tid < 256) S.homtab[tid] = 1.0 / (1.0 + (double)(tid * tid));
    if (tid < 34) S.dummy[tid] = 0u;
    if (gt < 2) { H.box[gt][0] = 1 << 30; H.box[gt][1] = -1; H.box[gt][2] = 1 << 30; H.box[gt][3] = -1; }
    if (gt < 2 * kMaxAngles * 8) (&H.acc[0][0][0])[gt] = 0u;
    __syncthreads();
    // K1 (same stream, earlier launch) already wrote the tile maximum into the table when the basic
    // block is requested; then the max pass and its barrier are skipped.
    const bool k1_max = P.col_basic >= 0;
    // deferred epilogue of this group's previous tile
    bool prev_active = false;
    double* prev_row = nullptr;
    uint32_t* prev_status = nullptr;
    int prev_slot = 0, prev_h = 0, prev_w = 0;
    auto deferred_epilogue = [&](int pbuf) {
        if (prev_active && gw < P.n_angles && lane == 0) {
            unsigned long long hom_sum = 0ull;
            for (int w = 0; w < R.gwarps; ++w) hom_sum += H.whom[pbuf][gw][w];
            k3_epilogue<MASKED>(P, prev_row, prev_status, prev_slot, prev_h, prev_w, gw, H.acc[pbuf][gw], hom_sum);
#pragma unroll
            for (int k = 0; k < 8; ++k) H.acc[pbuf][gw][k] = 0u;
        }
        if (MASKED && prev_active && gt == 0) {
            H.box[pbuf][0] = 1 << 30; H.box[pbuf][1] = -1; H.box[pbuf][2] = 1 << 30; H.box[pbuf][3] = -1;
        }
        prev_active = false;
    };

    const long long first = blockIdx.x;
    const long long mine = first < P.n_tiles ? (P.n_tiles - first + gridDim.x - 1) / gridDim.x : 0;
    const long long n_iter = (mine + ng - 1) / ng;
    TileWalk walk;
    walk.init(P, first + (long long)g * gridDim.x < P.n_tiles ? first + (long long)g * gridDim.x : 0,
              (long long)ng * gridDim.x);
    for (long long it = 0; it < n_iter; ++it, walk.next()) {
        const int buf = (int)(it & 1);
        const long long kk = (long long)ng * it + g;
        const bool active = kk < mine;
        const long long t = first + kk * gridDim.x;
        Tile T;
        T.h = 0; T.w = 0; T.n = 0;
        if (active) {
            T = resolve_tile_rs(P, walk.row, walk.slot);
            const uint4* px4 = reinterpret_cast<const uint4*>(T.px);
            const uint2* mk2 = reinterpret_cast<const uint2*>(T.mk);
            const int nfull = T.n >> 3, rem = T.n & 7;
            uint8_t* mbytes = reinterpret_cast<uint8_t*>(Gp.mbits);

            // ---- 1. tile maximum (over the mask when masked); stage the mask bits and their bounding box ----
            uint32_t mx2 = 0u;
            int brmin = 1 << 30, brmax = -1, bcmin = 1 << 30, bcmax = -1;
            double vmaxd = 0.0;
            if (k1_max) vmaxd = T.out_row[P.col_basic + kNBasic * T.slot + 10];
            for (int idx = gt; idx < nfull && (MASKED || !k1_max); idx += gthreads) {
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (!k1_max) v = ld_reuse(px4 + idx);
                if (MASKED) {
                    const uint2 m = __ldg(mk2 + idx);
                    const uint32_t c0 = __vcmpne4(m.x, 0u), c1 = __vcmpne4(m.y, 0u);
                    // 8 mask bytes -> 8 bits (byte k -> bit k)
                    const uint32_t b0 = ((c0 & 0x01010101u) * 0x01020408u) >> 24;
                    const uint32_t b1 = ((c1 & 0x01010101u) * 0x01020408u) >> 24;
                    const uint32_t bits8 = (b0 & 0xfu) | ((b1 & 0xfu) << 4);
                    mbytes[idx] = (uint8_t)bits8;
                    if (bits8) {
                        const int p0 = 8 * idx, ra = p0 / T.w, ca = p0 - ra * T.w;
                        if (ca + 7 < T.w) {                // the 8 pixels lie in one row
                            brmin = min(brmin, ra); brmax = max(brmax, ra);
                            bcmin = min(bcmin, ca + __ffs(bits8) - 1); bcmax = max(bcmax, ca + 31 - __clz(bits8));
                        } else {                           // straddles rows: be conservative
                            brmin = min(brmin, ra); brmax = max(brmax, (p0 + 7) / T.w);
                            bcmin = 0; bcmax = T.w - 1;
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
                    const bool ok = !MASKED || T.mk[i] != 0;
                    if (ok) {
                        bits |= 1u << k;
                        if (!k1_max) mx2 = __vmaxu2(mx2, (uint32_t)T.px[i]);
                        const int ra = i / T.w, ca = i - ra * T.w;
                        brmin = min(brmin, ra); brmax = max(brmax, ra); bcmin = min(bcmin, ca); bcmax = max(bcmax, ca);
                    }
                }
                if (MASKED) mbytes[nfull] = (uint8_t)bits;
            }
            if (MASKED) {
                brmin = __reduce_min_sync(0xffffffffu, brmin); brmax = __reduce_max_sync(0xffffffffu, brmax);
                bcmin = __reduce_min_sync(0xffffffffu, bcmin); bcmax = __reduce_max_sync(0xffffffffu, bcmax);
                if (lane == 0 && brmax >= 0) {
                    atomicMin(&H.box[buf][0], brmin); atomicMax(&H.box[buf][1], brmax);
                    atomicMin(&H.box[buf][2], bcmin); atomicMax(&H.box[buf][3], bcmax);
                }
            }
            uint32_t vmax;
            if (k1_max) {
                vmax = (vmaxd == vmaxd) ? (uint32_t)vmaxd : 0u;      // NaN: empty mask, no pair exists anyway
            } else {
                const uint32_t wm = __reduce_max_sync(0xffffffffu, max(mx2 & 0xffffu, mx2 >> 16));
                if (lane == 0) H.wmax[gw] = wm;
                ring_group_sync(R);
                vmax = lane < R.gwarps ? H.wmax[lane] : 0u;
                vmax = __reduce_max_sync(0xffffffffu, vmax);
            }

            // ---- 2. quantise to 8 bits into shared memory ----
            uint32_t mul = 0, sh = 24;
            if (lane == 0) k3_magic(vmax, mul, sh);
            mul = __shfl_sync(0xffffffffu, mul, 0);
            sh = __shfl_sync(0xffffffffu, sh, 0);
            for (int idx = gt; idx < nfull; idx += gthreads) {
                const uint4 v = ld_reuse(px4 + idx);
                const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
                uint32_t q[2] = {0u, 0u};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    // pixels outside the mask may exceed vmax; they never enter a pair, clamp them
                    uint32_t a = k3_quant(w4[k] & 0xffffu, mul, sh), b = k3_quant(w4[k] >> 16, mul, sh);
                    if (MASKED) { a = min(a, 255u); b = min(b, 255u); }
                    q[k >> 1] |= (a | (b << 8)) << (16 * (k & 1));
                }
                *reinterpret_cast<uint2*>(Gp.q8 + 2 * idx) = make_uint2(q[0], q[1]);
            }
            if (gt < rem) {
                const int i = nfull * 8 + gt;
                reinterpret_cast<uint8_t*>(Gp.q8)[i] = (uint8_t)min(k3_quant(T.px[i], mul, sh), 255u);
            }
        }
        ring_group_sync(R);                                // this tile staged; the previous one is complete
        deferred_epilogue(buf ^ 1);

        // ---- 3. one GLCM per direction ----
        // box of pixels that can take part in a pair: the tile, or the mask's bounding box
        int bx[4] = {0, T.h - 1, 0, T.w - 1};
        if (MASKED && active) { bx[0] = H.box[buf][0]; bx[1] = H.box[buf][1]; bx[2] = H.box[buf][2]; bx[3] = H.box[buf][3]; }
        for (int a = 0; a < P.n_angles; ++a) {
            const K3Geom G = k3_geom(T.w, P.dr[a], P.dc[a], bx[0], bx[1], bx[2], bx[3]);
            uint32_t I4[kK3Cache], J4[kK3Cache], vm[kK3Cache];
            if (active) {
                // table-free: pair-stream sums; the first kK3Cache items of a thread stay in registers
                K3Acc A = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0.0};
#pragma unroll
                for (int i = 0; i < kK3Cache; ++i) {
                    const int item = gt + i * gthreads;
                    vm[i] = 0u; I4[i] = 0u; J4[i] = 0u;
                    if (item < G.items && k3_item<MASKED>(Gp, G, item, I4[i], J4[i], vm[i]))
                        k3_sums<MASKED>(S, I4[i], J4[i], vm[i], A);
                }
                for (int item = gt + kK3Cache * gthreads; item < G.items; item += gthreads) {
                    uint32_t i4, j4, v;
                    if (k3_item<MASKED>(Gp, G, item, i4, j4, v)) k3_sums<MASKED>(S, i4, j4, v, A);
                }
                uint32_t red[6] = {A.si, A.sj, A.sii, A.sjj, A.sij, A.sd};
#pragma unroll
                for (int k = 0; k < 6; ++k) red[k] = __reduce_add_sync(0xffffffffu, red[k]);
                const uint32_t mm = __reduce_add_sync(0xffffffffu, A.m);
                // per-thread double -> 2^-40 fixed point, then exact integer sums (order independent)
                const unsigned long long hf = warp_sum_redux((unsigned long long)__double2ll_rn(A.hom * 1099511627776.0));
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) atomicAdd(&H.acc[buf][a][k], red[k]);
                    atomicAdd(&H.acc[buf][a][7], mm);
                    H.whom[buf][a][gw] = hf;
                }
            }
            ring_acquire(R);                               // ---- table owned by this group ----
            if (active) {
                uint32_t sold = 0u;
#pragma unroll
                for (int i = 0; i < kK3Cache; ++i)
                    if (vm[i]) k3_bins<0, MASKED>(S, I4[i], J4[i], vm[i], sold);
                for (int item = gt + kK3Cache * gthreads; item < G.items; item += gthreads) {
                    uint32_t i4, j4, v;
                    if (k3_item<MASKED>(Gp, G, item, i4, j4, v)) k3_bins<0, MASKED>(S, i4, j4, v, sold);
                }
                ring_group_sync(R);                        // bins complete
                if (DUMP) {
                    uint32_t* dst = P.counts + (t * P.n_angles + a) * 65536ll;
                    for (int k = gt; k < 32768; k += gthreads) {
                        const uint32_t wv = S.hist[k];
                        reinterpret_cast<uint2*>(dst)[k] = make_uint2(wv & 0xffffu, wv >> 16);
                    }
                    ring_group_sync(R);
                }
                uint32_t dummy = 0u;
#pragma unroll
                for (int i = 0; i < kK3Cache; ++i)
                    if (vm[i]) k3_bins<2, MASKED>(S, I4[i], J4[i], vm[i], dummy);
                for (int item = gt + kK3Cache * gthreads; item < G.items; item += gthreads) {
                    uint32_t i4, j4, v;
                    if (k3_item<MASKED>(Gp, G, item, i4, j4, v)) k3_bins<2, MASKED>(S, i4, j4, v, dummy);
                }
                ring_release(R);                           // ---- hand the table to the next group ----
                sold = __reduce_add_sync(0xffffffffu, sold);
                if (lane == 0) atomicAdd(&H.acc[buf][a][6], sold);
            } else {
                ring_release(R);
            }
        }

        // ---- 4. the epilogue is deferred to the next round (after its staging barrier) ----
        prev_active = active;
        if (active) {
            prev_row = T.out_row; prev_status = T.status; prev_slot = T.slot; prev_h = T.h; prev_w = T.w;
        }
    }
    ring_group_sync(R);                                    // last tile of this group complete
    deferred_epilogue((int)(n_iter & 1) ^ 1);
}

}  // namespace imfeat
