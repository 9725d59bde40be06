// aux_kernels.cuh -- layout conversion and the synthetic-object generator.
#pragma once
#include "common.cuh"

namespace imfeat {

// README.md:8-9 stores an object as image (h,w,c) 16-bit and mask (h,w,c).  Convert a padded
// interleaved batch uint16[N][hs][ws][c] to the plane-compact planar layout.  One thread per
// valid pixel reads its c interleaved samples (contiguous) and scatters them to c planes
// (each store coalesced across the warp).
// mask_bits != 0: mhwc is bit-packed, bit k of an object = element k of its padded (hs, ws, c) block, every
// object starting at a multiple of 8 bytes (include/imfeat.h, IMFEAT_MASK_BITS_BYTES).
__global__ void pack_hwc_kernel(const uint16_t* __restrict__ hwc, const uint8_t* __restrict__ mhwc,
                                const int32_t* __restrict__ sizes, long long n_objects, int c,
                                int hs, int ws, long long plane_stride,
                                uint16_t* __restrict__ planes, uint8_t* __restrict__ masks, int mask_bits) {
    const long long per_obj = (long long)hs * ws;
    const long long obj_bits_bytes = ((per_obj * c + 63) / 64) * 8;
    const long long total = n_objects * per_obj;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const long long obj = g / per_obj;
        const int pix = (int)(g - obj * per_obj);
        const int r = pix / ws, col = pix - r * ws;
        const int h = sizes ? sizes[2 * obj] : hs, w = sizes ? sizes[2 * obj + 1] : ws;
        if (r >= h || col >= w) continue;
        const long long src = (obj * per_obj + pix) * c;
        const long long dst = obj * c * plane_stride + (long long)r * w + col;
        for (int ch = 0; ch < c; ++ch) {
            planes[dst + ch * plane_stride] = hwc[src + ch];
            if (masks) {
                if (mask_bits) {
                    const long long k = (long long)pix * c + ch;
                    masks[dst + ch * plane_stride] = (mhwc[obj * obj_bits_bytes + (k >> 3)] >> (k & 7)) & 1;
                } else {
                    masks[dst + ch * plane_stride] = mhwc[src + ch];
                }
            }
        }
    }
}

// Bit-packed planar masks -> byte masks: one thread per 8 elements (one packed byte -> one 64-bit store).
__global__ void unpack_mask_bits_kernel(const uint8_t* __restrict__ bits, long long n_planes, long long plane_stride,
                                        uint8_t* __restrict__ masks) {
    const long long plane_bits_bytes = ((plane_stride + 63) / 64) * 8, groups = plane_stride >> 3;   // plane_stride % 8 == 0
    const long long total = n_planes * groups;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const long long pl = g / groups, k = g - pl * groups;
        const uint32_t b = bits[pl * plane_bits_bytes + k];
        const uint32_t lo = ((b & 0xfu) * 0x00204081u) & 0x01010101u, hi = (((b >> 4) & 0xfu) * 0x00204081u) & 0x01010101u;
        reinterpret_cast<uint2*>(masks + pl * plane_stride)[k] = make_uint2(lo, hi);
    }
}

// ---- counter-based synthetic objects; numpy mirror: <package>/synth.py (keep in sync) ----
__host__ __device__ __forceinline__ unsigned long long sm64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void synth_kernel(unsigned long long seed, long long first_object, long long n_objects,
                             int c, int hs, int ws, long long plane_stride, int variable, int hmin,
                             int wmin, int mask_shrink, uint16_t* __restrict__ planes,
                             uint8_t* __restrict__ masks, int32_t* __restrict__ sizes) {
    // one CTA per (object, channel) plane
    for (long long pl = blockIdx.x; pl < n_objects * c; pl += gridDim.x) {
        const long long obj = pl / c;
        const int ch = (int)(pl - obj * c);
        const unsigned long long okey = sm64(seed ^ sm64((unsigned long long)(first_object + obj)));
        int h = hs, w = ws;
        if (variable) {
            h = hmin + (int)((okey & 0xffffull) % (unsigned)(hs - hmin + 1));
            w = wmin + (int)(((okey >> 16) & 0xffffull) % (unsigned)(ws - wmin + 1));
        }
        if (sizes && ch == 0 && threadIdx.x == 0) { sizes[2 * obj] = h; sizes[2 * obj + 1] = w; }
        const unsigned long long k = sm64(okey + (unsigned long long)(ch + 1) * 0xD1B54A32D192ED03ull);
        const unsigned long long k1 = sm64(k ^ 1ull), k2 = sm64(k ^ 2ull);
        const int offset = 100 + (int)((k1 & 0xffffull) % 901ull);
        const int sig = 8 + (int)(((k1 >> 16) & 0xffull) % 25ull);
        const long long amp = 500 + (long long)(((k1 >> 24) & 0xffffull) % 2596ull);
        const int fx = 56 + (int)((k2 & 0xffull) % 56ull), fy = 56 + (int)(((k2 >> 8) & 0xffull) % 56ull);
        const long long RX = max(2, (2 * w * fx) >> 8), RY = max(2, (2 * h * fy) >> 8);
        const int cx2 = (w - 1) + (int)(((k2 >> 16) & 0xffull) % (unsigned)(w / 4 + 1)) - w / 8;
        const int cy2 = (h - 1) + (int)(((k2 >> 24) & 0xffull) % (unsigned)(h / 4 + 1)) - h / 8;
        const long long D = RX * RX * RY * RY;
        uint16_t* dst = planes + pl * plane_stride;
        uint8_t* mdst = masks ? masks + pl * plane_stride : nullptr;
        for (int idx = threadIdx.x; idx < h * w; idx += blockDim.x) {
            const int r = idx / w, col = idx - r * w;
            const long long dx = 2 * col - cx2, dy = 2 * r - cy2;
            const long long E = dx * dx * RY * RY + dy * dy * RX * RX;
            const long long blob = (E < D) ? amp * (D - E) / D : 0;
            const unsigned long long u = sm64(k + 0x1000ull + (unsigned long long)idx);
            const int s4 = (int)(u & 0xff) + (int)((u >> 8) & 0xff) + (int)((u >> 16) & 0xff) +
                           (int)((u >> 24) & 0xff);
            const int noise = ((s4 - 510) * sig) >> 7;
            long long val = offset + noise + blob;
            val = val < 0 ? 0 : (val > 65535 ? 65535 : val);
            dst[idx] = (uint16_t)val;
            if (mdst) mdst[idx] = (E * 256 < D * mask_shrink) ? 1 : 0;
        }
    }
}

}  // namespace imfeat
