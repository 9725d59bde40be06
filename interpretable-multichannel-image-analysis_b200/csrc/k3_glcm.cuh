// k3_glcm.cuh -- K3: gray-level quantisation + GLCM + Haralick properties, one CTA per tile.
//
// Replaces, per channel:
//   (x / x.max()) * 255 -> uint8                      NB:293-295
//   greycomatrix(q, [5], [0], levels=256)              NB:298   (not symmetric, not normed)
//   greycoprops x6 (contrast .. correlation)           NB:301-306
//
// * quantiser: floor(255*x/max) with an exact multiply-shift reciprocal (bit-identical to the
//   notebook's float64 expression for every uint16 pair; tests/test_quantiser.py).
// * the 256x256 bins live in shared memory as 16-bit counters (two per 32-bit word, 128 KB),
//   built with shared-memory atomics on the pair stream, dumped on request (parity), and only
//   ever cleared sparsely by re-walking the pairs.
// * contrast / dissimilarity / correlation come from exact integer sums over the pair stream,
//   four pairs per SIMD video instruction (dp4a, vabsdiff4); ASM = sum_bins c^2 is evaluated
//   as sum_pairs c[bin(pair)], so no pass over the 65,536 bins is ever made.
#pragma once
#include "common.cuh"

namespace imfeat {

constexpr int kK3Threads = 512;
constexpr int kK3Warps = kK3Threads / 32;

struct K3Smem {
    uint32_t hist[32768];
    double homtab[256];
    double whom[kK3Warps];
    uint32_t wred[kK3Warps][8];
    uint32_t wmax[kK3Warps];
    uint32_t q8[kMaxPixels / 4 + 4];   // quantised pixels, bytes, + slack for unaligned reads
    uint32_t m8[kMaxPixels / 4 + 4];   // 0xff / 0x00 per pixel (masked variant only)
};

struct K3Acc {
    uint32_t si, sj, sii, sjj, sij, sd, sasm, m;
    double hom;
};

// exact floor(255*x / vmax) for 0 <= x <= vmax <= 65535:  (255*x * mul) >> sh
__device__ __forceinline__ void k3_magic(uint32_t vmax, uint32_t& mul, uint32_t& sh) {
    if (vmax == 0) { mul = 0; sh = 24; return; }
    const uint32_t l = (vmax <= 1) ? 0u : 32u - (uint32_t)__clz(vmax - 1);   // ceil(log2 vmax)
    sh = 24u + l;
    mul = (uint32_t)ceil(ldexp(1.0, (int)sh) / (double)vmax);
}
__device__ __forceinline__ uint32_t k3_quant(uint32_t x, uint32_t mul, uint32_t sh) {
    return (uint32_t)(((unsigned long long)(x * 255u) * mul) >> sh);
}

// four consecutive bytes starting at byte offset off (any alignment)
__device__ __forceinline__ uint32_t k3_load4(const uint32_t* b, int off) {
    const int w = off >> 2;
    return __funnelshift_r(b[w], b[w + 1], (off & 3) << 3);
}

template <int PHASE, bool MASKED>
__device__ __forceinline__ void k3_pass(K3Smem& S, int h, int w, int dr, int dc, K3Acc& A) {
    const int nrows = h - dr;                      // dr >= 0 for all supported directions
    const int c0 = dc < 0 ? -dc : 0;
    const int c1 = dc < 0 ? w : w - dc;
    if (nrows <= 0 || c1 <= c0) return;
    const int gpr = (c1 - c0 + 3) >> 2;            // groups of 4 pairs per row
    const int lg = gpr <= 1 ? 0 : 32 - __clz(gpr - 1);   // ceil(log2 gpr): row index = item >> lg
    const int items = nrows << lg;
    for (int item = threadIdx.x; item < items; item += kK3Threads) {
        const int r = item >> lg, cg = item & ((1 << lg) - 1);
        if (cg >= gpr) continue;
        const int c = c0 + 4 * cg;
        const int valid = min(4, c1 - c);
        const int oi = r * w + c, oj = oi + dr * w + dc;
        uint32_t vm = valid == 4 ? 0xffffffffu : ((1u << (8 * valid)) - 1u);
        if (MASKED) vm &= k3_load4(S.m8, oi) & k3_load4(S.m8, oj);
        const uint32_t I4 = k3_load4(S.q8, oi) & vm, J4 = k3_load4(S.q8, oj) & vm;
        if (PHASE == 0) {
            A.si = __dp4a(I4, 0x01010101u, A.si);
            A.sj = __dp4a(J4, 0x01010101u, A.sj);
            A.sii = __dp4a(I4, I4, A.sii);
            A.sjj = __dp4a(J4, J4, A.sjj);
            A.sij = __dp4a(I4, J4, A.sij);
            A.sd += __vsadu4(I4, J4);
            if (MASKED) A.m += __popc(vm) >> 3;
        }
        const uint32_t D4 = __vabsdiffu4(I4, J4);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (!((vm >> (8 * b)) & 1u)) continue;
            const uint32_t key = (((I4 >> (8 * b)) & 0xffu) << 8) | ((J4 >> (8 * b)) & 0xffu);
            if (PHASE == 0) {
                A.hom += S.homtab[(D4 >> (8 * b)) & 0xffu];
                atomicAdd(&S.hist[key >> 1], 1u << ((key & 1u) << 4));
            }
            if (PHASE == 1) A.sasm += (S.hist[key >> 1] >> ((key & 1u) << 4)) & 0xffffu;
            if (PHASE == 2) S.hist[key >> 1] = 0u;
        }
    }
}

template <bool MASKED, bool DUMP>
__global__ void __launch_bounds__(kK3Threads, 1) k3_glcm_kernel(const __grid_constant__ Params P) {
    extern __shared__ __align__(16) unsigned char k3_smem_raw[];
    K3Smem& S = *reinterpret_cast<K3Smem*>(k3_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int k = tid; k < 32768; k += kK3Threads) S.hist[k] = 0u;
    if (tid < 256) S.homtab[tid] = 1.0 / (1.0 + (double)(tid * tid));
    __syncthreads();

    for (long long t = blockIdx.x; t < P.n_tiles; t += gridDim.x) {
        const Tile T = resolve_tile(P, t);
        const uint4* px4 = reinterpret_cast<const uint4*>(T.px);
        const uint2* mk2 = reinterpret_cast<const uint2*>(T.mk);
        const int nfull = T.n >> 3, rem = T.n & 7;

        // ---- 1. tile maximum (over the mask when masked); stage the mask bytes ----
        uint32_t mx2 = 0u;
        for (int idx = tid; idx < nfull; idx += kK3Threads) {
            uint4 v = ld_reuse(px4 + idx);
            if (MASKED) {
                const uint2 m = __ldg(mk2 + idx);
                const uint32_t c0 = __vcmpne4(m.x, 0u), c1 = __vcmpne4(m.y, 0u);
                S.m8[2 * idx] = c0;
                S.m8[2 * idx + 1] = c1;
                v.x &= __byte_perm(c0, 0u, 0x1100); v.y &= __byte_perm(c0, 0u, 0x3322);
                v.z &= __byte_perm(c1, 0u, 0x1100); v.w &= __byte_perm(c1, 0u, 0x3322);
            }
            mx2 = __vmaxu2(mx2, __vmaxu2(__vmaxu2(v.x, v.y), __vmaxu2(v.z, v.w)));
        }
        if (tid < rem) {
            const int i = nfull * 8 + tid;
            const bool ok = !MASKED || T.mk[i] != 0;
            if (MASKED) reinterpret_cast<uint8_t*>(S.m8)[i] = ok ? 0xffu : 0u;
            if (ok) mx2 = __vmaxu2(mx2, (uint32_t)T.px[i]);
        }
        const uint32_t wm = __reduce_max_sync(0xffffffffu, max(mx2 & 0xffffu, mx2 >> 16));
        if (lane == 0) S.wmax[warp] = wm;
        __syncthreads();
        uint32_t vmax = 0;
#pragma unroll
        for (int k = 0; k < kK3Warps; ++k) vmax = max(vmax, S.wmax[k]);

        // ---- 2. quantise to 8 bits into shared memory ----
        uint32_t mul = 0, sh = 24;
        if (lane == 0) k3_magic(vmax, mul, sh);
        mul = __shfl_sync(0xffffffffu, mul, 0);
        sh = __shfl_sync(0xffffffffu, sh, 0);
        for (int idx = tid; idx < nfull; idx += kK3Threads) {
            const uint4 v = ld_reuse(px4 + idx);
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
            uint32_t q[2] = {0u, 0u};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                // pixels outside the mask may exceed vmax; they never enter a pair, clamp them
                const uint32_t a = min(k3_quant(w4[k] & 0xffffu, mul, sh), 255u);
                const uint32_t b = min(k3_quant(w4[k] >> 16, mul, sh), 255u);
                q[k >> 1] |= (a | (b << 8)) << (16 * (k & 1));
            }
            S.q8[2 * idx] = q[0];
            S.q8[2 * idx + 1] = q[1];
        }
        if (tid < rem) {
            const int i = nfull * 8 + tid;
            reinterpret_cast<uint8_t*>(S.q8)[i] = (uint8_t)min(k3_quant(T.px[i], mul, sh), 255u);
        }
        __syncthreads();

        // ---- 3. one GLCM per direction ----
        for (int a = 0; a < P.n_angles; ++a) {
            const int dr = P.dr[a], dc = P.dc[a];
            K3Acc A = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0.0};
            k3_pass<0, MASKED>(S, T.h, T.w, dr, dc, A);
            __syncthreads();                               // bins complete
            if (DUMP) {
                uint32_t* dst = P.counts + (t * P.n_angles + a) * 65536ll;
                for (int k = tid; k < 32768; k += kK3Threads) {
                    const uint32_t wv = S.hist[k];
                    reinterpret_cast<uint2*>(dst)[k] = make_uint2(wv & 0xffffu, wv >> 16);
                }
            }
            k3_pass<1, MASKED>(S, T.h, T.w, dr, dc, A);
            uint32_t red[8] = {A.si, A.sj, A.sii, A.sjj, A.sij, A.sd, A.sasm, A.m};
#pragma unroll
            for (int k = 0; k < 8; ++k) red[k] = __reduce_add_sync(0xffffffffu, red[k]);
            const double hom = warp_sum(A.hom);
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 8; ++k) S.wred[warp][k] = red[k];
                S.whom[warp] = hom;
            }
            __syncthreads();                               // all read-backs done, partials visible
            k3_pass<2, MASKED>(S, T.h, T.w, dr, dc, A);
            if (tid == 0) {
                unsigned long long s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                double homt = 0.0;
                for (int k = 0; k < kK3Warps; ++k) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) s[j] += S.wred[k][j];
                    homt += S.whom[k];
                }
                long long M;
                if (MASKED) M = (long long)s[7];
                else {
                    const int nrows = T.h - dr, c0 = dc < 0 ? -dc : 0, c1 = dc < 0 ? T.w : T.w - dc;
                    M = (nrows > 0 && c1 > c0) ? (long long)nrows * (c1 - c0) : 0;
                }
                double* o = T.out_row + P.col_glcm + (T.slot * P.n_angles + a) * kNGlcm;
                if (M == 0) {
                    o[0] = 0.0; o[1] = 0.0; o[2] = 0.0; o[3] = 0.0; o[4] = 0.0; o[5] = 1.0;
                    if (T.status) atomicOr(T.status, kStNoPairs);
                } else {
                    const double Md = (double)M;
                    const long long con = (long long)s[2] + (long long)s[3] - 2ll * (long long)s[4];
                    const long long vi = M * (long long)s[2] - (long long)s[0] * (long long)s[0];
                    const long long vj = M * (long long)s[3] - (long long)s[1] * (long long)s[1];
                    const long long cov = M * (long long)s[4] - (long long)s[0] * (long long)s[1];
                    const double asmv = (double)s[6] / (Md * Md);
                    o[0] = (double)con / Md;
                    o[1] = (double)s[5] / Md;
                    o[2] = homt / Md;
                    o[3] = asmv;
                    o[4] = sqrt(asmv);
                    o[5] = (vi == 0 || vj == 0) ? 1.0 : (double)cov / (sqrt((double)vi) * sqrt((double)vj));
                }
            }
            __syncthreads();                               // bins clean, partials consumed
        }
    }
}

}  // namespace imfeat
