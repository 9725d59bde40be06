// k4_shape.cuh -- K4: mask shape descriptors and intensity-weighted spatial moments, one CTA per
// tile.  EXTENSION: the reference has no counterpart (README.md:9 only documents the optional
// mask key); the specification is oracle/notebook_oracle.py::shape_values / moment_values.
//
//   shape   : area, perimeter (skimage.measure.perimeter, 4-neighbourhood: border = mask minus
//             its 4-connected erosion, classified by the weighted 3x3 border sum), bounding box,
//             extent, centroid, axis lengths and eccentricity of the inertia tensor, circularity.
//   moments : weighted centroid and the normalised central moments nu_pq, 2 <= p+q <= 3.
//
// Fast path (h, w <= 256): rows are bit-packed with __ballot_sync; erosion, border and the
// perimeter classes are evaluated 32 pixels per bitwise instruction (bit-sliced adders), and
// every sum -- including the raw intensity moments up to order 3 -- is an exact integer, turned
// into central moments with 128-bit arithmetic in the epilogue.  Extreme aspect ratios fall back
// to a simple per-pixel path.
#pragma once
#include "common.cuh"

namespace imfeat {

constexpr int kK4Threads = 256;
constexpr int kK4Warps = kK4Threads / 32;
constexpr int kK4FastDim = 256;        // fast path: h, w <= 256 and h*w <= kK4FastPixels
constexpr int kK4FastPixels = 16384;
constexpr int kK4FastWords = 1024;     // >= h * ceil(w/32) under the limits above (<= 768)
constexpr int kK4NInt = 22;   // integer partials per warp

struct K4Smem {
    unsigned long long wint[kK4Warps][kK4NInt];
    int wbox[kK4Warps][4];
    double wdbl[kK4Warps][7];
    union {
        struct {
            uint32_t mrow[kK4FastWords]; uint32_t brow[kK4FastWords];
            uint4 px16[kK4FastPixels / 8];      // the tile, staged with 128-bit loads
            uint2 m8[kK4FastPixels / 8];        // its mask bytes
        } fast;
        struct { uint8_t m8[kMaxPixels + 16]; uint8_t b8[kMaxPixels + 16]; } slow;
    } u;
};

__device__ __forceinline__ double i128_to_double(__int128 v) {
    const bool neg = v < 0;
    const unsigned __int128 a = neg ? (unsigned __int128)(-v) : (unsigned __int128)v;
    const double d = __ull2double_rn((unsigned long long)(a >> 64)) * 18446744073709551616.0 +
                     __ull2double_rn((unsigned long long)a);
    return neg ? -d : d;
}

// s: 0 area, 1 sr, 2 sc, 3 srr, 4 scc, 5 src, 6 n1, 7 n2, 8 n3
__device__ __forceinline__ void k4_shape_epilogue(const Params& P, const Tile& T,
                                                  const unsigned long long* s, int rmin, int rmax,
                                                  int cmin, int cmax) {
    double* o = T.out_row + P.col_shape + kNShape * T.slot;
    const double SQ2 = 1.4142135623730951;
    const double perim = (double)s[6] + (double)s[7] * SQ2 + (double)s[8] * ((1.0 + SQ2) / 2.0);
    o[0] = (double)s[0];
    o[1] = perim;
    if (s[0] == 0) {
        for (int k = 2; k < kNShape; ++k) o[k] = qnan();
        if (T.status) atomicOr(T.status, kStEmptyMask);
        return;
    }
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

// raw moments m[0..9] = M00 M10 M01 M20 M11 M02 M30 M21 M12 M03 (exact integers, r,c <= 255)
__device__ __forceinline__ void k4_moment_epilogue(const Params& P, const Tile& T, const unsigned long long* m) {
    double* o = T.out_row + P.col_moment + kNMoment * T.slot;
    if (m[0] == 0) {
        for (int k = 0; k < kNMoment; ++k) o[k] = qnan();
        return;
    }
    typedef __int128 i128;
    const i128 M00 = m[0], M10 = m[1], M01 = m[2], M20 = m[3], M11 = m[4], M02 = m[5];
    const i128 M30 = m[6], M21 = m[7], M12 = m[8], M03 = m[9];
    const double M = (double)m[0];
    // mu_pq * M00^(p+q-1), exactly
    const double e20 = i128_to_double(M20 * M00 - M10 * M10);
    const double e11 = i128_to_double(M11 * M00 - M10 * M01);
    const double e02 = i128_to_double(M02 * M00 - M01 * M01);
    const double e30 = i128_to_double(M30 * M00 * M00 - 3 * M10 * M20 * M00 + 2 * M10 * M10 * M10);
    const double e21 = i128_to_double(M21 * M00 * M00 - 2 * M10 * M11 * M00 - M01 * M20 * M00 + 2 * M10 * M10 * M01);
    const double e12 = i128_to_double(M12 * M00 * M00 - 2 * M01 * M11 * M00 - M10 * M02 * M00 + 2 * M01 * M01 * M10);
    const double e03 = i128_to_double(M03 * M00 * M00 - 3 * M01 * M02 * M00 + 2 * M01 * M01 * M01);
    const double M3 = M * M * M, M45 = M * M * pow(M, 2.5);   // M00^3, M00^4.5
    o[0] = (double)m[1] / M;
    o[1] = (double)m[2] / M;
    o[2] = e20 / M3; o[3] = e11 / M3; o[4] = e02 / M3;
    o[5] = e30 / M45; o[6] = e21 / M45; o[7] = e12 / M45; o[8] = e03 / M45;
}

template <bool MASKED>
__global__ void __launch_bounds__(kK4Threads, 3) k4_shape_kernel(const __grid_constant__ Params P) {
    extern __shared__ __align__(16) unsigned char k4_smem_raw[];
    K4Smem& S = *reinterpret_cast<K4Smem*>(k4_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool want_mom = P.col_moment >= 0;

    TileWalk walk;
    walk.init(P, blockIdx.x < P.n_tiles ? blockIdx.x : 0, gridDim.x);
    for (long long t = blockIdx.x; t < P.n_tiles; t += gridDim.x, walk.next()) {
        const Tile T = resolve_tile_rs(P, walk.row, walk.slot);
        const int h = T.h, w = T.w, n = T.n;
        unsigned long long a[kK4NInt];
#pragma unroll
        for (int k = 0; k < kK4NInt; ++k) a[k] = 0ull;
        // a: 0 area 1 sr 2 sc 3 srr 4 scc 5 src 6 n1 7 n2 8 n3 | 9.. raw moments M00 M10 M01 M20 M11
        //    M02 M30 M21 M12 M03 (fast path) or M00 M10 M01 (slow path)
        int rmin = 1 << 30, rmax = -1, cmin = 1 << 30, cmax = -1;
        const bool fast = h <= kK4FastDim && w <= kK4FastDim && n <= kK4FastPixels;

        if (fast) {
            const int Pw = (w + 31) >> 5;                  // mask words per row
            // ---- stage the tile (and its mask) in shared memory: all loads of a thread in flight ----
            {
                const uint4* px4 = reinterpret_cast<const uint4*>(T.px);
                const uint2* mk2 = reinterpret_cast<const uint2*>(T.mk);
                const int nv = (n + 7) >> 3;
                for (int i = tid; i < nv; i += kK4Threads) {
                    if (want_mom) S.u.fast.px16[i] = ld_stream(px4 + i);
                    if (MASKED) S.u.fast.m8[i] = __ldg(mk2 + i);
                }
            }
            __syncthreads();
            const uint16_t* spx = reinterpret_cast<const uint16_t*>(S.u.fast.px16);
            const uint8_t* smk = reinterpret_cast<const uint8_t*>(S.u.fast.m8);
            // ---- pass A: warps over rows, lanes over columns.  Row sums are accumulated per lane
            //      for one 32-column block at a time and folded with the (lane-constant) column
            //      index afterwards: M_pq = sum_c c^q * (sum_r r^p * I[r][c]) ----
            uint32_t area = 0, sr = 0, sc = 0, srr = 0, scc = 0, src = 0;
            unsigned long long mq[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // M00 M10 M01 M20 M11 M02 M30 M21 M12 M03
            for (int c0 = 0; c0 < w; c0 += 32) {
                const int c = c0 + lane;
                const bool inb = c < w;
                uint32_t cnt = 0, csr = 0, csrr = 0, b0 = 0, b1 = 0;
                unsigned long long b2 = 0, b3 = 0;
                // running pointers: row `warp` of this column block, advanced by kK4Warps rows per step
                const uint8_t* mp = smk + warp * w + (inb ? c : 0);
                const uint16_t* pp = spx + warp * w + (inb ? c : 0);
                uint32_t* mr = S.u.fast.mrow + warp * Pw + (c0 >> 5);
                const int mstep = kK4Warps * w, wstep = kK4Warps * Pw;
#pragma unroll 4
                for (int r = warp; r < h; r += kK4Warps) {
                    const bool m = inb && (!MASKED || *mp != 0);
                    const uint32_t pxv = want_mom ? (uint32_t)*pp : 0u;
                    const uint32_t bal = __ballot_sync(0xffffffffu, m);
                    if (lane == 0) *mr = bal;
                    const uint32_t m1 = m ? 1u : 0u, im = m ? pxv : 0u;
                    const uint32_t r1 = r, r2 = r * r, r3 = r2 * r;
                    cnt += m1; csr += m1 * r1; csrr += m1 * r2;
                    rmin = m ? min(rmin, r) : rmin; rmax = m ? max(rmax, r) : rmax;
                    b0 += im; b1 += im * r1;
                    b2 += (unsigned long long)im * r2;
                    b3 += (unsigned long long)im * r3;
                    mp += mstep; pp += mstep; mr += wstep;
                }
                if (cnt) { cmin = min(cmin, c); cmax = max(cmax, c); }
                const uint32_t c1 = inb ? c : 0, c2 = c1 * c1, c3 = c2 * c1;
                area += cnt; sr += csr; srr += csrr; sc += c1 * cnt; scc += c2 * cnt; src += c1 * csr;
                if (want_mom) {
                    mq[0] += b0; mq[1] += b1; mq[3] += b2; mq[6] += b3;
                    mq[2] += (unsigned long long)b0 * c1; mq[4] += (unsigned long long)b1 * c1;
                    mq[7] += b2 * c1;
                    mq[5] += (unsigned long long)b0 * c2; mq[8] += (unsigned long long)b1 * c2;
                    mq[9] += (unsigned long long)b0 * c3;
                }
            }
            a[0] = __reduce_add_sync(0xffffffffu, area); a[1] = __reduce_add_sync(0xffffffffu, sr);
            a[2] = __reduce_add_sync(0xffffffffu, sc);   a[3] = __reduce_add_sync(0xffffffffu, srr);
            a[4] = __reduce_add_sync(0xffffffffu, scc);  a[5] = __reduce_add_sync(0xffffffffu, src);
            if (want_mom) {
#pragma unroll
                for (int k = 0; k < 10; ++k) a[9 + k] = warp_sum_redux(mq[k]);
            }
            __syncthreads();                               // mask rows complete
            // ---- pass B: border = mask & ~erosion4(mask), 32 pixels per word ----
            const int words = h * Pw;
            for (int idx = tid; idx < words; idx += kK4Threads) {
                const int r = idx / Pw, cw = idx - r * Pw;
                const uint32_t* M = S.u.fast.mrow;
                const uint32_t m = M[idx];
                const uint32_t up = r > 0 ? M[idx - Pw] : 0u, dn = r + 1 < h ? M[idx + Pw] : 0u;
                const uint32_t ml = cw > 0 ? M[idx - 1] : 0u, mr = cw + 1 < Pw ? M[idx + 1] : 0u;
                const uint32_t L = (m << 1) | (ml >> 31), R = (m >> 1) | (mr << 31);
                S.u.fast.brow[idx] = m & ~(up & dn & L & R);
            }
            __syncthreads();
            // ---- pass C: classify border pixels by their 3x3 border neighbourhood ----
            uint32_t n1 = 0, n2 = 0, n3 = 0;
            for (int idx = tid; idx < words; idx += kK4Threads) {
                const uint32_t* B = S.u.fast.brow;
                const uint32_t b = B[idx];
                if (!b) continue;
                const int r = idx / Pw, cw = idx - r * Pw;
                const bool hu = r > 0, hd = r + 1 < h, hl = cw > 0, hr = cw + 1 < Pw;
                const uint32_t bl = hl ? B[idx - 1] : 0u, br = hr ? B[idx + 1] : 0u;
                const uint32_t u = hu ? B[idx - Pw] : 0u, ul = (hu && hl) ? B[idx - Pw - 1] : 0u;
                const uint32_t ur = (hu && hr) ? B[idx - Pw + 1] : 0u;
                const uint32_t d = hd ? B[idx + Pw] : 0u, dl = (hd && hl) ? B[idx + Pw - 1] : 0u;
                const uint32_t dr = (hd && hr) ? B[idx + Pw + 1] : 0u;
                const uint32_t N = u, Sd = d, W = (b << 1) | (bl >> 31), E = (b >> 1) | (br << 31);
                const uint32_t NW = (u << 1) | (ul >> 31), NE = (u >> 1) | (ur << 31);
                const uint32_t SW = (d << 1) | (dl >> 31), SE = (d >> 1) | (dr << 31);
                // bit-sliced sums o = N+S+W+E, dd = NW+NE+SW+SE (0..4 each)
                uint32_t s1 = N ^ Sd ^ W, c1 = (N & Sd) | (N & W) | (Sd & W);
                const uint32_t o0 = s1 ^ E, c2 = s1 & E, o1 = c1 ^ c2, o2 = c1 & c2;
                s1 = NW ^ NE ^ SW; c1 = (NW & NE) | (NW & SW) | (NE & SW);
                const uint32_t d0 = s1 ^ SE, c3 = s1 & SE, d1 = c1 ^ c3, d2 = c1 & c3;
                const uint32_t o_is0 = ~o0 & ~o1 & ~o2, o_is1 = o0 & ~o1 & ~o2, o_23 = o1 & ~o2;
                const uint32_t d_is1 = d0 & ~d1 & ~d2, d_is2 = ~d0 & d1 & ~d2, d_is3 = d0 & d1 & ~d2;
                const uint32_t d_le2 = ~d2 & ~(d1 & d0);
                n1 += __popc(b & o_23 & d_le2);                        // weights {5,7,15,17,25,27}
                n2 += __popc(b & ((o_is0 & d_is2) | (o_is1 & d_is3))); // {21,33}
                n3 += __popc(b & o_is1 & (d_is1 | d_is2));             // {13,23}
            }
            a[6] = __reduce_add_sync(0xffffffffu, n1);
            a[7] = __reduce_add_sync(0xffffffffu, n2);
            a[8] = __reduce_add_sync(0xffffffffu, n3);
        } else {
            // ---- generic per-pixel path (very wide / very tall tiles) ----
            for (int i = tid; i < n; i += kK4Threads) S.u.slow.m8[i] = MASKED ? (T.mk[i] != 0) : 1;
            __syncthreads();
            for (int i = tid; i < n; i += kK4Threads) {
                const int r = i / w, c = i - r * w;
                uint8_t b = 0;
                if (S.u.slow.m8[i]) {
                    const bool up = r > 0 && S.u.slow.m8[i - w], dn = r + 1 < h && S.u.slow.m8[i + w];
                    const bool lf = c > 0 && S.u.slow.m8[i - 1], rt = c + 1 < w && S.u.slow.m8[i + 1];
                    b = !(up && dn && lf && rt);
                }
                S.u.slow.b8[i] = b;
            }
            __syncthreads();
            for (int i = tid; i < n; i += kK4Threads) {
                if (!S.u.slow.m8[i]) continue;
                const int r = i / w, c = i - r * w;
                a[0] += 1; a[1] += r; a[2] += c;
                a[3] += (unsigned long long)r * r; a[4] += (unsigned long long)c * c;
                a[5] += (unsigned long long)r * c;
                rmin = min(rmin, r); rmax = max(rmax, r); cmin = min(cmin, c); cmax = max(cmax, c);
                if (S.u.slow.b8[i]) {
                    int v = 1;
#pragma unroll
                    for (int dr = -1; dr <= 1; ++dr)
#pragma unroll
                        for (int dc = -1; dc <= 1; ++dc) {
                            if (dr == 0 && dc == 0) continue;
                            const int rr = r + dr, cc = c + dc;
                            if (rr < 0 || rr >= h || cc < 0 || cc >= w) continue;
                            if (S.u.slow.b8[rr * w + cc]) v += (dr != 0 && dc != 0) ? 10 : 2;
                        }
                    const unsigned long long C1 = (1ull << 5) | (1ull << 7) | (1ull << 15) | (1ull << 17) |
                                                  (1ull << 25) | (1ull << 27);
                    const unsigned long long C2 = (1ull << 21) | (1ull << 33);
                    const unsigned long long C3 = (1ull << 13) | (1ull << 23);
                    a[6] += (C1 >> v) & 1ull; a[7] += (C2 >> v) & 1ull; a[8] += (C3 >> v) & 1ull;
                }
                if (want_mom) {
                    const unsigned long long x = T.px[i];
                    a[9] += x; a[10] += x * r; a[11] += x * c;
                }
            }
#pragma unroll
            for (int k = 0; k < 12; ++k) a[k] = warp_sum(a[k]);
        }
        rmin = __reduce_min_sync(0xffffffffu, rmin); rmax = __reduce_max_sync(0xffffffffu, rmax);
        cmin = __reduce_min_sync(0xffffffffu, cmin); cmax = __reduce_max_sync(0xffffffffu, cmax);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < kK4NInt; ++k) S.wint[warp][k] = a[k];
            S.wbox[warp][0] = rmin; S.wbox[warp][1] = rmax; S.wbox[warp][2] = cmin; S.wbox[warp][3] = cmax;
        }
        __syncthreads();
        unsigned long long s[kK4NInt];
        if (warp == 0 || !fast) {
#pragma unroll
            for (int k = 0; k < kK4NInt; ++k) {
                s[k] = 0;
                for (int wv = 0; wv < kK4Warps; ++wv) s[k] += S.wint[wv][k];
            }
            for (int wv = 0; wv < kK4Warps; ++wv) {
                rmin = min(rmin, S.wbox[wv][0]); rmax = max(rmax, S.wbox[wv][1]);
                cmin = min(cmin, S.wbox[wv][2]); cmax = max(cmax, S.wbox[wv][3]);
            }
        }
        if (tid == 0 && P.col_shape >= 0) k4_shape_epilogue(P, T, s, rmin, rmax, cmin, cmax);
        if (want_mom) {
            if (fast) {
                if (tid == 32 % kK4Threads) {
                    // second warp's lane 0 re-sums (keeps the two epilogues on different warps)
                    unsigned long long mm[10];
                    for (int k = 0; k < 10; ++k) {
                        mm[k] = 0;
                        for (int wv = 0; wv < kK4Warps; ++wv) mm[k] += S.wint[wv][9 + k];
                    }
                    k4_moment_epilogue(P, T, mm);
                }
            } else {
                // slow path: second pass around the weighted centroid in double precision
                double* o = T.out_row + P.col_moment + kNMoment * T.slot;
                if (s[9] == 0) {
                    if (tid == 0) for (int k = 0; k < kNMoment; ++k) o[k] = qnan();
                } else {
                    const double M = (double)s[9];
                    const double cr = (double)s[10] / M, cc = (double)s[11] / M;
                    double mu[7] = {0, 0, 0, 0, 0, 0, 0};   // 20 11 02 30 21 12 03
                    for (int i = tid; i < n; i += kK4Threads) {
                        if (!S.u.slow.m8[i]) continue;
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
        }
        __syncthreads();   // shared arrays are rewritten by the next tile
    }
}

}  // namespace imfeat

// -------------------------------------------------------------------------------------------------
// K4w: one WARP per tile (32-thread CTAs, many per SM) for batches whose stride fits the fast path
// (Hs, Ws <= 256, Hs*Ws <= 16384).  Same arithmetic as the fast path above, but with no CTA-wide
// barrier, no cross-warp reduction and no staging pass: rows are read straight from global memory
// (four rows in flight per lane) and latency is hidden by occupancy.
// -------------------------------------------------------------------------------------------------
namespace imfeat {

constexpr int kK4wWords = 1024;     // >= h * ceil(w/32) under the fast-path limits (<= 768)
constexpr int kK4wBatch = 16;       // finished tiles parked per warp until their epilogues run side by side

// The two epilogues (eigenvalues, square roots, 128-bit central moments: ~700 instructions) were a fifth
// of the per-tile cost with one lane each; parked, 16 tiles are finished by 32 lanes at once.
struct K4Pending {
    double* out_row;
    uint32_t* status;
    int slot, rmin, rmax, cmin, cmax, pad;
    unsigned long long s[9];
    unsigned long long mq[10];
};

#ifndef IMFEAT_K4W_WARPS
#define IMFEAT_K4W_WARPS 20            // resident warps per SM the register budget is set for
#endif
// GENERAL: the batch has tiles whose rows are not 8, 16, .. 256 pixels long (a size table, or such a stride); the
// kernel then carries the general vector pass as well, which costs the other variant registers (and 4 % of its speed).
template <bool MASKED, bool GENERAL>
__global__ void __launch_bounds__(32, IMFEAT_K4W_WARPS) k4w_shape_kernel(const __grid_constant__ Params P) {
    __shared__ uint32_t mrow[kK4wWords];
    __shared__ uint32_t brow[kK4wWords];
    __shared__ K4Pending pending[kK4wBatch];
    const int lane = threadIdx.x;
    const bool want_mom = P.col_moment >= 0;
    int n_pending = 0;
    auto flush = [&](int count) {
        __syncwarp();
        const int k = lane & (kK4wBatch - 1);
        if (k < count) {
            const K4Pending& q = pending[k];
            Tile Tq;
            Tq.out_row = q.out_row; Tq.status = q.status; Tq.slot = q.slot;
            if (lane < kK4wBatch) { if (P.col_shape >= 0) k4_shape_epilogue(P, Tq, q.s, q.rmin, q.rmax, q.cmin, q.cmax); }
            else if (want_mom) k4_moment_epilogue(P, Tq, q.mq);
        }
        __syncwarp();
    };

    long long tnext = next_tile(P.sched + 3);
    while (tnext < P.n_tiles) {
        const long long t = tnext;
        tnext = next_tile(P.sched + 3);                    // one tile ahead
        const Tile T = resolve_tile(P, t);
        const int h = T.h, w = T.w;
        const int Pw = (w + 31) >> 5;
        int rmin = 1 << 30, rmax = -1, cmin = 1 << 30, cmax = -1;
        uint32_t area = 0, sr = 0, sc = 0, srr = 0, scc = 0, src = 0;
        unsigned long long mq[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // M00 M10 M01 M20 M11 M02 M30 M21 M12 M03

        // ---- pass A ----
        const int cpr = w >> 3;                            // 8-pixel chunks per row
        if ((w & 7) == 0 && (cpr & (cpr - 1)) == 0 && cpr <= 32) {
            // Vector pass (w = 8, 16, .. 256): chunks of 8 consecutive pixels lane-strided over the tile, one
            // 128-bit pixel load + one 64-bit mask load each.  32 is a multiple of cpr, so a lane keeps its
            // column c0 and walks down the rows: per chunk only the local sums sum_k k^q I_k (IDP.2A with
            // constant weights) and sum_k k^q m_k (IDP.4A) times r^p are accumulated; the column index is
            // folded in once per tile.
            const int lg = 31 - __clz(cpr), lc = lane & (cpr - 1), c0 = lc << 3;
            const int rstep = 32 >> lg, nchunk = h << lg;
            if (w & 31) {                                  // bytes behind the row end inside the last word
                for (int k = lane; k < h * Pw; k += 32) mrow[k] = 0u;
                __syncwarp();
            }
            uint8_t* mbytes = reinterpret_cast<uint8_t*>(mrow) + lc;
            const uint4* px4 = reinterpret_cast<const uint4*>(T.px);
            const uint2* mk2 = reinterpret_cast<const uint2*>(T.mk);
            uint32_t cntL = 0, skL = 0, sk2L = 0, rc1 = 0, rc2 = 0, rsk = 0, colbits = 0;
            uint32_t B00 = 0, B01 = 0, B02 = 0, B03 = 0;
            unsigned long long B10 = 0, B20 = 0, B30 = 0, B11 = 0, B21 = 0, B12 = 0;
            auto chunk = [&](int r, uint4 v, uint2 m) {
                uint32_t bits8 = 0xffu, b03 = 0x01010101u, b47 = 0x01010101u;
                if (MASKED && (m.x | m.y) == 0u) {          // nothing of the mask here: only its (zero) bits are recorded
                    mbytes[r * (Pw << 2)] = (uint8_t)0;
                    return;
                }
                if (MASKED) {
                    const uint32_t n0 = __vcmpne4(m.x, 0u), n1 = __vcmpne4(m.y, 0u);
                    b03 = n0 & 0x01010101u; b47 = n1 & 0x01010101u;
                    bits8 = (((b03 * 0x01020408u) >> 24) & 0xfu) | (((b47 * 0x01020408u) >> 20) & 0xf0u);
                    v.x &= __byte_perm(n0, 0u, 0x1100); v.y &= __byte_perm(n0, 0u, 0x3322);
                    v.z &= __byte_perm(n1, 0u, 0x1100); v.w &= __byte_perm(n1, 0u, 0x3322);
                }
                mbytes[r * (Pw << 2)] = (uint8_t)bits8;
                const uint32_t cnt = __popc(bits8);
                const uint32_t sk = __dp4a(b47, 0x07060504u, __dp4a(b03, 0x03020100u, 0u));
                const uint32_t sk2 = __dp4a(b47, 0x31241910u, __dp4a(b03, 0x09040100u, 0u));
                const uint32_t r1 = (uint32_t)r, r2 = r1 * r1;
                cntL += cnt; skL += sk; sk2L += sk2; rc1 += r1 * cnt; rc2 += r2 * cnt; rsk += r1 * sk;
                colbits |= bits8;
                if (cnt) { rmin = min(rmin, r); rmax = r; }
                if (want_mom) {
                    const uint32_t s0 = __dp2a_lo(v.w, 0x0101u, __dp2a_lo(v.z, 0x0101u, __dp2a_lo(v.y, 0x0101u, __dp2a_lo(v.x, 0x0101u, 0u))));
                    const uint32_t s1 = __dp2a_lo(v.w, 0x0706u, __dp2a_lo(v.z, 0x0504u, __dp2a_lo(v.y, 0x0302u, __dp2a_lo(v.x, 0x0100u, 0u))));
                    const uint32_t s2 = __dp2a_lo(v.w, 0x3124u, __dp2a_lo(v.z, 0x1910u, __dp2a_lo(v.y, 0x0904u, __dp2a_lo(v.x, 0x0100u, 0u))));
                    // cubes 0 1 8 27 64 125 216 343: 343 = 255 + 88 does not fit one byte weight
                    const uint32_t s3 = __dp2a_lo(v.w, 0x5800u, __dp2a_lo(v.w, 0xffd8u, __dp2a_lo(v.z, 0x7d40u, __dp2a_lo(v.y, 0x1b08u, __dp2a_lo(v.x, 0x0100u, 0u)))));
                    const uint32_t r3 = r2 * r1;
                    B00 += s0; B01 += s1; B02 += s2; B03 += s3;
                    B10 += (unsigned long long)s0 * r1; B20 += (unsigned long long)s0 * r2; B30 += (unsigned long long)s0 * r3;
                    B11 += (unsigned long long)s1 * r1; B21 += (unsigned long long)s1 * r2;
                    B12 += (unsigned long long)s2 * r1;
                }
            };
            int idx = lane, r = lane >> lg;
            for (; idx + 96 < nchunk; idx += 128, r += 4 * rstep) {      // four chunks of loads in flight
                uint4 v[4];
                uint2 m[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    v[u] = want_mom ? ld_stream(px4 + idx + 32 * u) : make_uint4(0u, 0u, 0u, 0u);
                    m[u] = MASKED ? __ldg(mk2 + idx + 32 * u) : make_uint2(0u, 0u);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) chunk(r + u * rstep, v[u], m[u]);
            }
            for (; idx < nchunk; idx += 32, r += rstep)
                chunk(r, want_mom ? ld_stream(px4 + idx) : make_uint4(0u, 0u, 0u, 0u), MASKED ? __ldg(mk2 + idx) : make_uint2(0u, 0u));
            // fold the lane's column in: c = c0 + k
            const uint32_t c1 = (uint32_t)c0, c2 = c1 * c1;
            area = cntL; sr = rc1; srr = rc2;
            sc = c1 * cntL + skL; scc = c2 * cntL + 2u * c1 * skL + sk2L; src = c1 * rc1 + rsk;
            if (colbits) { cmin = c0 + __ffs(colbits) - 1; cmax = c0 + 31 - __clz(colbits); }
            if (want_mom) {
                const unsigned long long C1 = c1, C2 = c2, C3 = (unsigned long long)c2 * c1;
                mq[0] = B00; mq[1] = B10; mq[3] = B20; mq[6] = B30;
                mq[2] = C1 * B00 + B01; mq[4] = C1 * B10 + B11; mq[7] = C1 * B20 + B21;
                mq[5] = C2 * B00 + 2ull * C1 * B01 + B02; mq[8] = C2 * B10 + 2ull * C1 * B11 + B12;
                mq[9] = C3 * B00 + 3ull * C2 * B01 + 3ull * C1 * B02 + B03;
            }
        } else if (GENERAL && w >= 8) {
            // General vector pass (any w >= 8): 8-pixel chunks of the compact plane, lane-strided, the same 128-bit
            // pixel / 64-bit mask loads.  A chunk starts at (r, c) = divmod(8 * idx, w) and may run over the end of
            // its row (once: w >= 8): it is then worked off as two segments, (r, c) with its first w - c pixels and
            // (r + 1, c - w) with the others -- the column of pixel k is c + k in both, with a negative c in the
            // second.  The column is folded into the local sums per segment (it changes from chunk to chunk), all
            // in wrapping integer arithmetic whose true values are non-negative and fit.
            for (int k = lane; k < h * Pw; k += 32) mrow[k] = 0u;
            __syncwarp();
            const int n = T.n, nchunk = (n + 7) >> 3;
            const float rtw = __frcp_rn((float)w);
            const uint4* px4 = reinterpret_cast<const uint4*>(T.px);
            const uint2* mk2 = reinterpret_cast<const uint2*>(T.mk);
            auto seg = [&](int r, int c, uint4 v, uint32_t n0, uint32_t n1) {      // n0 / n1: 0xff per pixel of the segment
                const uint32_t b03 = n0 & 0x01010101u, b47 = n1 & 0x01010101u;
                const uint32_t bits8 = (((b03 * 0x01020408u) >> 24) & 0xfu) | (((b47 * 0x01020408u) >> 20) & 0xf0u);
                if (bits8 == 0u) return;
                {
                    const uint32_t bb = c < 0 ? bits8 >> (-c) : bits8;
                    const int pos = max(c, 0), sh = pos & 31;
                    IMFEAT_CHECK(r >= 0 && r < h && r * Pw + (pos >> 5) < kK4wWords && pos + 7 - (c < 0 ? -c : 0) < 32 * Pw + 8);
                    uint32_t* wp = mrow + r * Pw + (pos >> 5);
                    atomicOr(wp, bb << sh);
                    if (sh > 24 && (bb >> (32 - sh)) != 0u) atomicOr(wp + 1, bb >> (32 - sh));
                }
                const uint32_t cnt = __popc(bits8);
                const uint32_t sk = __dp4a(b47, 0x07060504u, __dp4a(b03, 0x03020100u, 0u));
                const uint32_t sk2 = __dp4a(b47, 0x31241910u, __dp4a(b03, 0x09040100u, 0u));
                const uint32_t r1 = (uint32_t)r, r2 = r1 * r1, cu = (uint32_t)c;
                const uint32_t colsum = cu * cnt + sk;
                area += cnt; sr += r1 * cnt; srr += r2 * cnt;
                sc += colsum; src += r1 * colsum; scc += cu * cu * cnt + 2u * cu * sk + sk2;
                rmin = min(rmin, r); rmax = max(rmax, r);
                cmin = min(cmin, c + __ffs(bits8) - 1); cmax = max(cmax, c + 31 - __clz(bits8));
                if (want_mom) {
                    v.x &= __byte_perm(n0, 0u, 0x1100); v.y &= __byte_perm(n0, 0u, 0x3322);
                    v.z &= __byte_perm(n1, 0u, 0x1100); v.w &= __byte_perm(n1, 0u, 0x3322);
                    const uint32_t s0 = __dp2a_lo(v.w, 0x0101u, __dp2a_lo(v.z, 0x0101u, __dp2a_lo(v.y, 0x0101u, __dp2a_lo(v.x, 0x0101u, 0u))));
                    const uint32_t s1 = __dp2a_lo(v.w, 0x0706u, __dp2a_lo(v.z, 0x0504u, __dp2a_lo(v.y, 0x0302u, __dp2a_lo(v.x, 0x0100u, 0u))));
                    const uint32_t s2 = __dp2a_lo(v.w, 0x3124u, __dp2a_lo(v.z, 0x1910u, __dp2a_lo(v.y, 0x0904u, __dp2a_lo(v.x, 0x0100u, 0u))));
                    const uint32_t s3 = __dp2a_lo(v.w, 0x5800u, __dp2a_lo(v.w, 0xffd8u, __dp2a_lo(v.z, 0x7d40u, __dp2a_lo(v.y, 0x1b08u, __dp2a_lo(v.x, 0x0100u, 0u)))));
                    // sum_k I_k (c + k)^q, q = 0..3 (64-bit wrapping; the true values are >= 0)
                    const unsigned long long C1 = (unsigned long long)(long long)c, C2 = C1 * C1;
                    const unsigned long long t0 = s0, t1 = C1 * s0 + s1, t2 = C2 * s0 + 2ull * C1 * s1 + s2;
                    const unsigned long long t3 = C2 * C1 * s0 + 3ull * C2 * s1 + 3ull * C1 * s2 + s3;
                    const unsigned long long R1 = r1, R2 = r2, R3 = (unsigned long long)r2 * r1;
                    mq[0] += t0; mq[1] += t0 * R1; mq[3] += t0 * R2; mq[6] += t0 * R3;
                    mq[2] += t1; mq[4] += t1 * R1; mq[7] += t1 * R2;
                    mq[5] += t2; mq[8] += t2 * R1;
                    mq[9] += t3;
                }
            };
            auto chunk = [&](int idx, const uint4& v, const uint2& m) {
                if (MASKED && (m.x | m.y) == 0u) return;   // nothing of the mask here (its row words are zero already)
                const int p0 = idx << 3;
                const int r = (int)(((float)p0 + 0.5f) * rtw), c = p0 - r * w;          // exact: p0 < 2^20
                const int nv = min(8, n - p0), na = min(nv, w - c);                    // pixels of the chunk; of them in row r
                // 0xff per pixel k < na (A) and na <= k < nv (B)
                const unsigned long long all = nv >= 8 ? ~0ull : ((1ull << (8 * nv)) - 1ull);
                const unsigned long long ma = na >= 8 ? ~0ull : ((1ull << (8 * na)) - 1ull);
                uint32_t n0 = 0xffffffffu, n1 = 0xffffffffu;
                if (MASKED) { n0 = __vcmpne4(m.x, 0u); n1 = __vcmpne4(m.y, 0u); }
                seg(r, c, v, n0 & (uint32_t)ma, n1 & (uint32_t)(ma >> 32));
                if (na < nv) seg(r + 1, c - w, v, n0 & (uint32_t)(all & ~ma), n1 & (uint32_t)((all & ~ma) >> 32));
            };
            int idx = lane;
            for (; idx + 96 < nchunk; idx += 128) {                               // four chunks of loads in flight
                uint4 v[4];
                uint2 m[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    v[u] = want_mom ? ld_stream(px4 + idx + 32 * u) : make_uint4(0u, 0u, 0u, 0u);
                    m[u] = MASKED ? __ldg(mk2 + idx + 32 * u) : make_uint2(0u, 0u);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) chunk(idx + 32 * u, v[u], m[u]);
            }
            for (; idx < nchunk; idx += 32)
                chunk(idx, want_mom ? ld_stream(px4 + idx) : make_uint4(0u, 0u, 0u, 0u), MASKED ? __ldg(mk2 + idx) : make_uint2(0u, 0u));
        } else {
            // ---- pass A: lanes over columns, rows in sequence (four rows of loads in flight) ----
            for (int c0 = 0; c0 < w; c0 += 32) {
                const int c = c0 + lane;
                const bool inb = c < w;
                uint32_t cnt = 0, csr = 0, csrr = 0, b0 = 0, b1 = 0;
                unsigned long long b2 = 0, b3 = 0;
                const uint8_t* mp = T.mk + (inb ? c : 0);
                const uint16_t* pp = T.px + (inb ? c : 0);
                uint32_t* mr = mrow + (c0 >> 5);
                auto row = [&](int r, bool m, uint32_t pxv) {
                    const uint32_t bal = __ballot_sync(0xffffffffu, m);
                    if (lane == 0) mr[r * Pw] = bal;
                    const uint32_t m1 = m ? 1u : 0u, im = m ? pxv : 0u;
                    const uint32_t r1 = r, r2 = r * r, r3 = r2 * r;
                    cnt += m1; csr += m1 * r1; csrr += m1 * r2;
                    rmin = m ? min(rmin, r) : rmin; rmax = m ? max(rmax, r) : rmax;
                    b0 += im; b1 += im * r1;
                    b2 += (unsigned long long)im * r2;
                    b3 += (unsigned long long)im * r3;
                };
                int r = 0;
                for (; r + 4 <= h; r += 4) {
                    bool m[4];
                    uint32_t pxv[4];
    #pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        m[u] = inb && (!MASKED || mp[(r + u) * w] != 0);
                        pxv[u] = (want_mom && inb) ? (uint32_t)pp[(r + u) * w] : 0u;
                    }
    #pragma unroll
                    for (int u = 0; u < 4; ++u) row(r + u, m[u], pxv[u]);
                }
                for (; r < h; ++r) {
                    const bool m = inb && (!MASKED || mp[r * w] != 0);
                    row(r, m, (want_mom && inb) ? (uint32_t)pp[r * w] : 0u);
                }
                if (cnt) { cmin = min(cmin, c); cmax = max(cmax, c); }
                const uint32_t c1 = inb ? c : 0, c2 = c1 * c1, c3 = c2 * c1;
                area += cnt; sr += csr; srr += csrr; sc += c1 * cnt; scc += c2 * cnt; src += c1 * csr;
                if (want_mom) {
                    mq[0] += b0; mq[1] += b1; mq[3] += b2; mq[6] += b3;
                    mq[2] += (unsigned long long)b0 * c1; mq[4] += (unsigned long long)b1 * c1;
                    mq[7] += b2 * c1;
                    mq[5] += (unsigned long long)b0 * c2; mq[8] += (unsigned long long)b1 * c2;
                    mq[9] += (unsigned long long)b0 * c3;
                }
            }
        }
        __syncwarp();
        // ---- pass B: border = mask & ~erosion4(mask), 32 pixels per word ----
        const int words = h * Pw;
        const float rPw = __frcp_rn((float)Pw);
        for (int idx = lane; idx < words; idx += 32) {
            const int r = (int)(((float)idx + 0.5f) * rPw), cw = idx - r * Pw;   // exact: idx < 2^20
            const uint32_t m = mrow[idx];
            const uint32_t up = r > 0 ? mrow[idx - Pw] : 0u, dn = r + 1 < h ? mrow[idx + Pw] : 0u;
            const uint32_t ml = cw > 0 ? mrow[idx - 1] : 0u, mr2 = cw + 1 < Pw ? mrow[idx + 1] : 0u;
            const uint32_t L = (m << 1) | (ml >> 31), Rr = (m >> 1) | (mr2 << 31);
            brow[idx] = m & ~(up & dn & L & Rr);
        }
        __syncwarp();
        // ---- pass C: classify border pixels by their 3x3 border neighbourhood ----
        uint32_t n1 = 0, n2 = 0, n3 = 0;
        for (int idx = lane; idx < words; idx += 32) {
            const uint32_t b = brow[idx];
            if (!b) continue;
            const int r = (int)(((float)idx + 0.5f) * rPw), cw = idx - r * Pw;   // exact: idx < 2^20
            const bool hu = r > 0, hd = r + 1 < h, hl = cw > 0, hr = cw + 1 < Pw;
            const uint32_t bl = hl ? brow[idx - 1] : 0u, br = hr ? brow[idx + 1] : 0u;
            const uint32_t u = hu ? brow[idx - Pw] : 0u, ul = (hu && hl) ? brow[idx - Pw - 1] : 0u;
            const uint32_t ur = (hu && hr) ? brow[idx - Pw + 1] : 0u;
            const uint32_t d = hd ? brow[idx + Pw] : 0u, dl = (hd && hl) ? brow[idx + Pw - 1] : 0u;
            const uint32_t dr = (hd && hr) ? brow[idx + Pw + 1] : 0u;
            const uint32_t N = u, Sd = d, W = (b << 1) | (bl >> 31), E = (b >> 1) | (br << 31);
            const uint32_t NW = (u << 1) | (ul >> 31), NE = (u >> 1) | (ur << 31);
            const uint32_t SW = (d << 1) | (dl >> 31), SE = (d >> 1) | (dr << 31);
            uint32_t s1 = N ^ Sd ^ W, c1 = (N & Sd) | (N & W) | (Sd & W);
            const uint32_t o0 = s1 ^ E, c2 = s1 & E, o1 = c1 ^ c2, o2 = c1 & c2;
            s1 = NW ^ NE ^ SW; c1 = (NW & NE) | (NW & SW) | (NE & SW);
            const uint32_t d0 = s1 ^ SE, c3 = s1 & SE, d1 = c1 ^ c3, d2 = c1 & c3;
            const uint32_t o_is0 = ~o0 & ~o1 & ~o2, o_is1 = o0 & ~o1 & ~o2, o_23 = o1 & ~o2;
            const uint32_t d_is1 = d0 & ~d1 & ~d2, d_is2 = ~d0 & d1 & ~d2, d_is3 = d0 & d1 & ~d2;
            const uint32_t d_le2 = ~d2 & ~(d1 & d0);
            n1 += __popc(b & o_23 & d_le2);
            n2 += __popc(b & ((o_is0 & d_is2) | (o_is1 & d_is3)));
            n3 += __popc(b & o_is1 & (d_is1 | d_is2));
        }
        // ---- warp totals (every lane gets them) and the two epilogues on two lanes ----
        unsigned long long s[9];
        s[0] = __reduce_add_sync(0xffffffffu, area); s[1] = __reduce_add_sync(0xffffffffu, sr);
        s[2] = __reduce_add_sync(0xffffffffu, sc);   s[3] = __reduce_add_sync(0xffffffffu, srr);
        s[4] = __reduce_add_sync(0xffffffffu, scc);  s[5] = __reduce_add_sync(0xffffffffu, src);
        s[6] = __reduce_add_sync(0xffffffffu, n1);   s[7] = __reduce_add_sync(0xffffffffu, n2);
        s[8] = __reduce_add_sync(0xffffffffu, n3);
        rmin = __reduce_min_sync(0xffffffffu, rmin); rmax = __reduce_max_sync(0xffffffffu, rmax);
        cmin = __reduce_min_sync(0xffffffffu, cmin); cmax = __reduce_max_sync(0xffffffffu, cmax);
        if (want_mom) {
#pragma unroll
            for (int k = 0; k < 10; ++k) mq[k] = warp_sum_redux(mq[k]);
        }
        if (lane == 0) {
            K4Pending& q = pending[n_pending];
            q.out_row = T.out_row; q.status = T.status; q.slot = T.slot;
            q.rmin = rmin; q.rmax = rmax; q.cmin = cmin; q.cmax = cmax;
#pragma unroll
            for (int k = 0; k < 9; ++k) q.s[k] = s[k];
            if (want_mom) {
#pragma unroll
                for (int k = 0; k < 10; ++k) q.mq[k] = mq[k];
            }
        }
        if (++n_pending == kK4wBatch) { flush(kK4wBatch); n_pending = 0; }
        __syncwarp();                                        // mrow / brow are rewritten by the next tile
    }
    flush(n_pending);
}

}  // namespace imfeat
