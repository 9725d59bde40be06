// k4_shape.cuh -- K4: mask shape descriptors and intensity-weighted spatial moments, one CTA per
// tile.  EXTENSION: the reference has no counterpart (README.md:9 only documents the optional
// mask key); the specification is oracle/notebook_oracle.py::shape_values / moment_values.
//
//   shape   : area, perimeter (skimage.measure.perimeter, 4-neighbourhood: border = mask minus
//             its 4-connected erosion, classified by the weighted 3x3 border sum), bounding box,
//             extent, centroid, axis lengths and eccentricity of the inertia tensor, circularity.
//             Every input of those formulas is an exact integer sum.
//   moments : weighted centroid and the normalised central moments nu_pq, 2 <= p+q <= 3.
#pragma once
#include "common.cuh"

namespace imfeat {

constexpr int kK4Threads = 256;
constexpr int kK4Warps = kK4Threads / 32;

struct K4Smem {
    uint8_t m8[kMaxPixels + 16];
    uint8_t b8[kMaxPixels + 16];
    unsigned long long wint[kK4Warps][12];
    int wbox[kK4Warps][4];
    double wdbl[kK4Warps][7];
    double centroid[2];
};

template <bool MASKED>
__global__ void __launch_bounds__(kK4Threads) k4_shape_kernel(const __grid_constant__ Params P) {
    extern __shared__ __align__(16) unsigned char k4_smem_raw[];
    K4Smem& S = *reinterpret_cast<K4Smem*>(k4_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (long long t = blockIdx.x; t < P.n_tiles; t += gridDim.x) {
        const Tile T = resolve_tile(P, t);
        const int h = T.h, w = T.w, n = T.n;
        // ---- 1. mask bytes (0/1) into shared memory ----
        for (int i = tid; i < n; i += kK4Threads) S.m8[i] = MASKED ? (T.mk[i] != 0) : 1;
        __syncthreads();
        // ---- 2. border image ----
        for (int i = tid; i < n; i += kK4Threads) {
            const int r = i / w, c = i - r * w;
            uint8_t b = 0;
            if (S.m8[i]) {
                const bool up = r > 0 && S.m8[i - w], dn = r + 1 < h && S.m8[i + w];
                const bool lf = c > 0 && S.m8[i - 1], rt = c + 1 < w && S.m8[i + 1];
                b = !(up && dn && lf && rt);
            }
            S.b8[i] = b;
        }
        __syncthreads();
        // ---- 3. integer sums ----
        unsigned long long a[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        // a: 0 area, 1 sr, 2 sc, 3 srr, 4 scc, 5 src, 6 n1, 7 n2, 8 n3, 9 M00, 10 M10, 11 M01
        int rmin = 1 << 30, rmax = -1, cmin = 1 << 30, cmax = -1;
        for (int i = tid; i < n; i += kK4Threads) {
            if (!S.m8[i]) continue;
            const int r = i / w, c = i - r * w;
            a[0] += 1; a[1] += r; a[2] += c;
            a[3] += (unsigned long long)r * r; a[4] += (unsigned long long)c * c;
            a[5] += (unsigned long long)r * c;
            rmin = min(rmin, r); rmax = max(rmax, r); cmin = min(cmin, c); cmax = max(cmax, c);
            if (S.b8[i]) {
                int v = 1;
#pragma unroll
                for (int dr = -1; dr <= 1; ++dr)
#pragma unroll
                    for (int dc = -1; dc <= 1; ++dc) {
                        if (dr == 0 && dc == 0) continue;
                        const int rr = r + dr, cc = c + dc;
                        if (rr < 0 || rr >= h || cc < 0 || cc >= w) continue;
                        if (S.b8[rr * w + cc]) v += (dr != 0 && dc != 0) ? 10 : 2;
                    }
                const unsigned long long C1 = (1ull << 5) | (1ull << 7) | (1ull << 15) | (1ull << 17) |
                                              (1ull << 25) | (1ull << 27);
                const unsigned long long C2 = (1ull << 21) | (1ull << 33);
                const unsigned long long C3 = (1ull << 13) | (1ull << 23);
                a[6] += (C1 >> v) & 1ull; a[7] += (C2 >> v) & 1ull; a[8] += (C3 >> v) & 1ull;
            }
            if (P.col_moment >= 0) {
                const unsigned long long x = T.px[i];
                a[9] += x; a[10] += x * r; a[11] += x * c;
            }
        }
#pragma unroll
        for (int k = 0; k < 12; ++k) a[k] = warp_sum(a[k]);
        rmin = __reduce_min_sync(0xffffffffu, rmin); rmax = __reduce_max_sync(0xffffffffu, rmax);
        cmin = __reduce_min_sync(0xffffffffu, cmin); cmax = __reduce_max_sync(0xffffffffu, cmax);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 12; ++k) S.wint[warp][k] = a[k];
            S.wbox[warp][0] = rmin; S.wbox[warp][1] = rmax; S.wbox[warp][2] = cmin; S.wbox[warp][3] = cmax;
        }
        __syncthreads();
        unsigned long long s[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            s[k] = 0;
            for (int wv = 0; wv < kK4Warps; ++wv) s[k] += S.wint[wv][k];
        }
        for (int wv = 0; wv < kK4Warps; ++wv) {
            rmin = min(rmin, S.wbox[wv][0]); rmax = max(rmax, S.wbox[wv][1]);
            cmin = min(cmin, S.wbox[wv][2]); cmax = max(cmax, S.wbox[wv][3]);
        }
        // ---- 4. shape epilogue ----
        if (tid == 0 && P.col_shape >= 0) {
            double* o = T.out_row + P.col_shape + kNShape * T.slot;
            const double SQ2 = 1.4142135623730951;
            const double perim = (double)s[6] + (double)s[7] * SQ2 + (double)s[8] * ((1.0 + SQ2) / 2.0);
            o[0] = (double)s[0];
            o[1] = perim;
            if (s[0] == 0) {
                for (int k = 2; k < kNShape; ++k) o[k] = qnan();
                if (T.status) atomicOr(T.status, kStEmptyMask);
            } else {
                const long long A = (long long)s[0];
                const double Ad = (double)A, A2 = Ad * Ad;
                const double bbox = (double)(rmax - rmin + 1) * (double)(cmax - cmin + 1);
                const long long an = A * (long long)s[4] - (long long)s[2] * (long long)s[2];
                const long long cn = A * (long long)s[3] - (long long)s[1] * (long long)s[1];
                const long long bn = A * (long long)s[5] - (long long)s[1] * (long long)s[2];
                const double ia = (double)an / A2, ic = (double)cn / A2, ib = -(double)bn / A2;
                const double hd = (double)(an - cn) / A2 * 0.5;
                const double D = sqrt(__dadd_rn(__dmul_rn(hd, hd), __dmul_rn(ib, ib)));
                const double l1 = __dadd_rn(__dmul_rn(__dadd_rn(ia, ic), 0.5), D);
                double l2 = 0.0, ecc = 0.0;
                if (l1 > 0) {
                    l2 = __ddiv_rn(__dsub_rn(__dmul_rn(ia, ic), __dmul_rn(ib, ib)), l1);
                    if (l2 < 0) l2 = 0;
                    ecc = __ddiv_rn(__dmul_rn(2.0, D), l1);
                    ecc = sqrt(fmin(fmax(ecc, 0.0), 1.0));
                }
                o[2] = bbox;
                o[3] = Ad / bbox;
                o[4] = (double)s[1] / Ad;
                o[5] = (double)s[2] / Ad;
                o[6] = 4.0 * sqrt(l1);
                o[7] = 4.0 * sqrt(l2);
                o[8] = ecc;
                o[9] = perim > 0 ? 4.0 * 3.14159265358979323846 * Ad / (perim * perim) : qnan();
            }
        }
        // ---- 5. spatial moments (second pass around the exact weighted centroid) ----
        if (P.col_moment >= 0) {
            double* o = T.out_row + P.col_moment + kNMoment * T.slot;
            if (s[9] == 0) {
                if (tid == 0) for (int k = 0; k < kNMoment; ++k) o[k] = qnan();
            } else {
                const double M = (double)s[9];
                const double cr = (double)s[10] / M, cc = (double)s[11] / M;
                double mu[7] = {0, 0, 0, 0, 0, 0, 0};   // 20 11 02 30 21 12 03
                for (int i = tid; i < n; i += kK4Threads) {
                    if (!S.m8[i]) continue;
                    const int r = i / w, c = i - r * w;
                    const double x = (double)T.px[i], dr = (double)r - cr, dc = (double)c - cc;
                    const double rr = dr * dr, cc2 = dc * dc;
                    mu[0] += rr * x; mu[1] += dr * dc * x; mu[2] += cc2 * x;
                    mu[3] += rr * dr * x; mu[4] += rr * dc * x; mu[5] += dr * cc2 * x; mu[6] += cc2 * dc * x;
                }
#pragma unroll
                for (int k = 0; k < 7; ++k) mu[k] = warp_sum(mu[k]);
                if (lane == 0)
#pragma unroll
                    for (int k = 0; k < 7; ++k) S.wdbl[warp][k] = mu[k];
                __syncthreads();
                if (tid == 0) {
                    double tot[7];
                    for (int k = 0; k < 7; ++k) {
                        tot[k] = 0.0;
                        for (int wv = 0; wv < kK4Warps; ++wv) tot[k] += S.wdbl[wv][k];
                    }
                    const double n2 = M * M, n3 = pow(M, 2.5);
                    o[0] = cr; o[1] = cc;
                    o[2] = tot[0] / n2; o[3] = tot[1] / n2; o[4] = tot[2] / n2;
                    o[5] = tot[3] / n3; o[6] = tot[4] / n3; o[7] = tot[5] / n3; o[8] = tot[6] / n3;
                }
            }
        }
        __syncthreads();   // shared arrays are rewritten by the next tile
    }
}

}  // namespace imfeat
