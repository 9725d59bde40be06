// K3a: staging for the GLCM kernel -- tile maximum, 8-bit quantisation, mask bits and their bounding box.
//
// Replaces the reference's (x / x.max()) * 255 -> uint8 step
// (channel_importance_hand_crafted_features.ipynb cell 13, NB:293-295) and prepares what K3
// (k3_glcm.cuh) needs per tile as one contiguous record: header (pair box, size), quantised bytes,
// mask bits.  One warp per tile, no CTA barrier, dynamic tile scheduling; the record is assembled in
// shared memory and leaves with a single bulk copy (cp.async.bulk shared -> global), so K3 can pull
// it into its ring with a single bulk copy as well.
#pragma once
#include "k3_glcm.cuh"

namespace imfeat {

constexpr int kK3aThreads = 32;

__host__ __device__ inline size_t k3a_smem_bytes(int max_pixels, bool masked) {
    return k3_rec_bytes(max_pixels, masked);
}

template <bool MASKED>
__global__ void __launch_bounds__(kK3aThreads, 16)
k3a_glcm_stage_kernel(const __grid_constant__ Params P, unsigned char* __restrict__ recs, int max_pixels) {
    extern __shared__ __align__(16) unsigned char k3a_raw[];
    unsigned char* rec = k3a_raw;
    K3RecHdr& Hd = *reinterpret_cast<K3RecHdr*>(rec);
    const uint32_t rec_bytes = (uint32_t)k3_rec_bytes(max_pixels, MASKED);
    K3Group Gp;
    Gp.q8 = reinterpret_cast<uint32_t*>(rec + sizeof(K3RecHdr));
    Gp.mbits = Gp.q8 + k3_q8_words(max_pixels);
    const int lane = threadIdx.x;
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

    }
    if (lane == 0) bulk_wait_all();
}

}  // namespace imfeat
