// K3a: the four GLCM properties that are plain sums over the pair stream.
//
// Replaces (together with k3_glcm.cuh) the reference's greycomatrix + greycoprops call
// (channel_importance_hand_crafted_features.ipynb cell 13, NB:269-308).  contrast, dissimilarity,
// homogeneity and correlation are linear in the co-occurrence matrix, so they need no matrix at all:
//   contrast      = sum (i-j)^2 / M = (Sii + Sjj - 2 Sij) / M
//   dissimilarity = sum |i-j| / M
//   homogeneity   = sum 1/(1+(i-j)^2) / M         (256-entry table, 2^-40 fixed point => order-free)
//   correlation   = (M Sij - Si Sj) / sqrt((M Sii - Si^2)(M Sjj - Sj^2))   (exact integers)
// Only ASM / energy need the bins (kernel K3, the 128 KB table ring).
//
// One warp per tile, no CTA barrier, dynamic tile scheduling: the quantised tile lives in a per-warp
// shared-memory buffer, every direction is one lane-strided pass over its pair groups.
#pragma once
#include "k3_glcm.cuh"

namespace imfeat {

constexpr int kK3aThreads = 32;

__host__ __device__ inline size_t k3a_smem_bytes(int max_pixels, bool masked) {
    return 256 * sizeof(double) + k3_rec_bytes(max_pixels, masked);
}

template <bool MASKED>
__global__ void __launch_bounds__(kK3aThreads, 16)
k3a_glcm_sums_kernel(const __grid_constant__ Params P, unsigned char* __restrict__ recs, int max_pixels) {
    extern __shared__ __align__(16) unsigned char k3a_raw[];
    double* homtab = reinterpret_cast<double*>(k3a_raw);
    // the tile's record is assembled in shared memory (K3a works on it) and stored with one bulk copy
    unsigned char* rec = k3a_raw + 256 * sizeof(double);
    K3RecHdr& Hd = *reinterpret_cast<K3RecHdr*>(rec);
    const uint32_t rec_bytes = (uint32_t)k3_rec_bytes(max_pixels, MASKED);
    K3Group Gp;
    Gp.q8 = reinterpret_cast<uint32_t*>(rec + sizeof(K3RecHdr));
    Gp.mbits = Gp.q8 + k3_q8_words(max_pixels);
    const int lane = threadIdx.x;
    for (int k = lane; k < 256; k += 32) homtab[k] = 1.0 / (1.0 + (double)(k * k));
    __syncwarp();
    const bool k1_max = P.col_basic >= 0;      // K1 (earlier launch, same stream) wrote the tile maximum

    for (long long t = next_tile(P.sched + 2); t < P.n_tiles; t = next_tile(P.sched + 2)) {
        const Tile T = resolve_tile(P, t);
        if (lane == 0) bulk_wait_read();                   // the previous record has left shared memory
        __syncwarp();
        const uint4* px4 = reinterpret_cast<const uint4*>(T.px);
        const uint2* mk2 = reinterpret_cast<const uint2*>(T.mk);
        const int nfull = T.n >> 3, rem = T.n & 7;
        uint8_t* mbytes = reinterpret_cast<uint8_t*>(Gp.mbits);

        // ---- 1. tile maximum (over the mask when masked), mask bits, bounding box ----
        uint32_t mx2 = 0u;
        int brmin = 1 << 30, brmax = -1, bcmin = 1 << 30, bcmax = -1;
        double vmaxd = 0.0;
        if (k1_max) vmaxd = T.out_row[P.col_basic + kNBasic * T.slot + 10];
        for (int idx = lane; idx < nfull && (MASKED || !k1_max); idx += 32) {
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (!k1_max) v = ld_reuse(px4 + idx);
            if (MASKED) {
                const uint2 m = __ldg(mk2 + idx);
                const uint32_t c0 = __vcmpne4(m.x, 0u), c1 = __vcmpne4(m.y, 0u);
                const uint32_t b0 = ((c0 & 0x01010101u) * 0x01020408u) >> 24;
                const uint32_t b1 = ((c1 & 0x01010101u) * 0x01020408u) >> 24;
                const uint32_t bits8 = (b0 & 0xfu) | ((b1 & 0xfu) << 4);
                mbytes[idx] = (uint8_t)bits8;
                if (bits8) {
                    const int p0 = 8 * idx, ra = p0 / T.w, ca = p0 - ra * T.w;
                    if (ca + 7 < T.w) {                    // the 8 pixels lie in one row
                        brmin = min(brmin, ra); brmax = max(brmax, ra);
                        bcmin = min(bcmin, ca + __ffs(bits8) - 1); bcmax = max(bcmax, ca + 31 - __clz(bits8));
                    } else {                               // straddles rows: be conservative
                        brmin = min(brmin, ra); brmax = max(brmax, (p0 + 7) / T.w);
                        bcmin = 0; bcmax = T.w - 1;
                    }
                }
                v.x &= __byte_perm(c0, 0u, 0x1100); v.y &= __byte_perm(c0, 0u, 0x3322);
                v.z &= __byte_perm(c1, 0u, 0x1100); v.w &= __byte_perm(c1, 0u, 0x3322);
            }
            mx2 = __vmaxu2(mx2, __vmaxu2(__vmaxu2(v.x, v.y), __vmaxu2(v.z, v.w)));
        }
        if (lane == 0 && rem) {                            // tail pixels (< 8): one thread, in order
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
        int bx[4] = {0, T.h - 1, 0, T.w - 1};
        if (MASKED) {
            bx[0] = __reduce_min_sync(0xffffffffu, brmin); bx[1] = __reduce_max_sync(0xffffffffu, brmax);
            bx[2] = __reduce_min_sync(0xffffffffu, bcmin); bx[3] = __reduce_max_sync(0xffffffffu, bcmax);
        }
        uint32_t vmax;
        if (k1_max) vmax = (vmaxd == vmaxd) ? (uint32_t)vmaxd : 0u;      // NaN: empty mask, no pair exists anyway
        else vmax = __reduce_max_sync(0xffffffffu, max(mx2 & 0xffffu, mx2 >> 16));

        // ---- 2. quantise to 8 bits into shared memory ----
        uint32_t mul = 0, sh = 24;
        if (lane == 0) k3_magic(vmax, mul, sh);
        mul = __shfl_sync(0xffffffffu, mul, 0);
        sh = __shfl_sync(0xffffffffu, sh, 0);
        for (int idx = lane; idx < nfull; idx += 32) {
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
        if (lane < rem) {
            const int i = nfull * 8 + lane;
            reinterpret_cast<uint8_t*>(Gp.q8)[i] = (uint8_t)min(k3_quant(T.px[i], mul, sh), 255u);
        }
        if (lane == 0) {
            Hd.box[0] = bx[0]; Hd.box[1] = bx[1]; Hd.box[2] = bx[2]; Hd.box[3] = bx[3];
            Hd.h = T.h; Hd.w = T.w; Hd.pad[0] = 0; Hd.pad[1] = 0;
        }
        bulk_fence_smem();
        __syncwarp();
        if (lane == 0) bulk_s2g(recs + (size_t)t * rec_bytes, smem_addr(rec), rec_bytes);

        // ---- 3. one lane-strided pass over the pair groups per direction ----
        for (int a = 0; a < P.n_angles; ++a) {
            const K3Geom G = k3_geom(T.w, P.dr[a], P.dc[a], bx[0], bx[1], bx[2], bx[3]);
            K3Acc A0 = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0.0}, A1 = A0;   // two independent chains
            uint32_t nproc = 0u;                           // items this lane went through
            for (int item = lane; item < G.items; item += 32) {
                uint32_t I4[4], J4[4], pm;
                if (k3_item16<MASKED>(Gp, G, item, I4, J4, pm)) {
                    k3_sums(homtab, I4[0], J4[0], k3_expand4(pm), A0);
                    k3_sums(homtab, I4[1], J4[1], k3_expand4(pm >> 4), A1);
                    k3_sums(homtab, I4[2], J4[2], k3_expand4(pm >> 8), A0);
                    k3_sums(homtab, I4[3], J4[3], k3_expand4(pm >> 12), A1);
                    A0.m += __popc(pm);
                    ++nproc;
                }
            }
            const uint32_t si = __reduce_add_sync(0xffffffffu, A0.si + A1.si);
            const uint32_t sj = __reduce_add_sync(0xffffffffu, A0.sj + A1.sj);
            const uint32_t sii = __reduce_add_sync(0xffffffffu, A0.sii + A1.sii);
            const uint32_t sjj = __reduce_add_sync(0xffffffffu, A0.sjj + A1.sjj);
            const uint32_t sij = __reduce_add_sync(0xffffffffu, A0.sij + A1.sij);
            const uint32_t sd = __reduce_add_sync(0xffffffffu, A0.sd + A1.sd);
            const uint32_t mm = __reduce_add_sync(0xffffffffu, A0.m + A1.m);
            const uint32_t np = __reduce_add_sync(0xffffffffu, nproc);
            // per-lane double sums (fixed lane-strided order) are rounded to 2^-40 fixed point, so the
            // warp reduction is an integer sum and the result does not depend on which warp ran the tile
            const unsigned long long homfix =
                warp_sum_redux((unsigned long long)__double2ll_rn((A0.hom + A1.hom) * 1099511627776.0));
            if (lane == 0) {
                double* o = T.out_row + P.col_glcm + (T.slot * P.n_angles + a) * kNGlcm;
                const long long M = mm;
                if (M == 0) {
                    o[0] = 0.0; o[1] = 0.0; o[2] = 0.0; o[5] = 1.0;
                    if (T.status) atomicOr(T.status, kStNoPairs);
                } else {
                    // the pairs of the processed items that do not exist added exactly 1.0 each
                    const long long D = 16ll * np - M;
                    const double Md = (double)M;
                    const long long Si = si, Sj = sj, Sii = sii, Sjj = sjj, Sij = sij;
                    o[0] = (double)(Sii + Sjj - 2 * Sij) / Md;
                    o[1] = (double)sd / Md;
                    o[2] = ((double)(homfix - ((unsigned long long)D << 40)) * 9.094947017729282e-13) / Md;
                    const long long vi = M * Sii - Si * Si, vj = M * Sjj - Sj * Sj, cov = M * Sij - Si * Sj;
                    o[5] = (vi == 0 || vj == 0) ? 1.0 : (double)cov / (sqrt((double)vi) * sqrt((double)vj));
                }
            }
        }
        __syncwarp();                                      // q8 / mbits are rewritten by the next tile
    }
    if (lane == 0) bulk_wait_all();
}

}  // namespace imfeat
