/*
 * imfeat_ref.c -- plain-C CPU restatement of the reference hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  It is the fast checker used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline leg; nothing in the product
 * package links or calls it.
 *
 * It follows /root/reference/channel_importance_hand_crafted_features.ipynb
 * (NB:<n> = raw JSON line n):
 *   basic block  NB:241-262  min, np.percentile(q=0.1..0.9), max, sum, mean, std,
 *                            scipy kurtosis (Fisher, biased), scipy skew (biased),
 *                            skimage shannon_entropy (base 2)
 *   GLCM block   NB:293-306  (x/max)*255 -> uint8, greycomatrix(d=5, angle 0, 256
 *                            levels, not symmetric, not normed), six greycoprops
 *   table        NB:330-331  basic block for Ch1..ChC, then GLCM block for Ch1..ChC
 * and the library algorithms those lines call (numpy 2.3.5 _function_base_impl.py
 * _lerp/_quantile, numpy _methods.py _var, scipy 1.18.1 _stats_py.py _moment/skew/
 * kurtosis, scikit-image texture.py / _texture.pyx), as laid out in SURVEY.md
 * section 8 (a) and Appendix A.  It is validated against oracle/notebook_oracle.py and
 * the committed golden fixtures (tests/test_oracle_c.py).
 *
 * Extensions with no reference counterpart (masks, 4 GLCM directions, shape, spatial
 * moments) follow the specification functions in oracle/notebook_oracle.py.
 *
 * Build: see oracle/Makefile (gcc -O2 -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define N_BASIC 17
#define N_GLCM 6
#define N_SHAPE 10
#define N_MOM 9
#define LEVELS 256

static int cmp_u16(const void *a, const void *b) {
    uint16_t x = *(const uint16_t *)a, y = *(const uint16_t *)b;
    return (x > y) - (x < y);
}

/* numpy.percentile(method="linear") on a sorted uint16 array; q in percent. */
static double percentile_sorted(const uint16_t *s, long n, double q_percent) {
    double quant = q_percent / 100.0;
    double virt = (double)(n - 1) * quant;
    long lo, hi;
    double g;
    if (virt >= (double)(n - 1)) {
        lo = hi = n - 1;
        g = virt - floor(virt);
    } else {
        lo = (long)floor(virt);
        hi = lo + 1;
        g = virt - (double)lo;
    }
    double a = (double)s[lo], b = (double)s[hi];
    double diff = (double)(uint16_t)(s[hi] - s[lo]);
    volatile double t1 = diff * g;          /* volatile: forbid fused multiply-add */
    volatile double t2 = diff * (1.0 - g);
    return (g >= 0.5) ? (b - t2) : (a + t1);
}

/* 17 basic features of the multiset v[0..n) (v is clobbered: sorted in place). */
static void basic_block(uint16_t *v, long n, double *out) {
    int k;
    if (n == 0) {
        for (k = 0; k < N_BASIC; ++k) out[k] = NAN;
        return;
    }
    uint64_t sum = 0;
    long i;
    for (i = 0; i < n; ++i) sum += v[i];
    double mean = (double)sum / (double)n;
    long double s2 = 0, s3 = 0, s4 = 0;
    for (i = 0; i < n; ++i) {
        double d = (double)v[i] - mean;
        double d2 = d * d;
        s2 += d2;
        s3 += d2 * d;
        s4 += d2 * d2;
    }
    double m2 = (double)(s2 / n), m3 = (double)(s3 / n), m4 = (double)(s4 / n);
    qsort(v, (size_t)n, sizeof(uint16_t), cmp_u16);
    out[0] = v[0];
    for (k = 1; k <= 9; ++k) out[k] = percentile_sorted(v, n, (double)k / 10.0);
    out[10] = v[n - 1];
    out[11] = (double)sum;
    out[12] = mean;
    out[13] = sqrt(m2);
    /* scipy: NaN when m2 <= (eps*mean)^2, i.e. a constant channel */
    double thr = 2.220446049250313e-16 * mean;
    if (m2 <= thr * thr) {
        out[14] = NAN;
        out[15] = NAN;
    } else {
        out[14] = m4 / (m2 * m2) - 3.0;
        out[15] = m3 / pow(m2, 1.5);
    }
    /* entropy over distinct raw values, base 2 */
    long double H = 0;
    i = 0;
    while (i < n) {
        long j = i;
        while (j < n && v[j] == v[i]) ++j;
        double p = (double)(j - i) / (double)n;
        H -= (long double)(p * log(p));
        i = j;
    }
    out[16] = (double)(H / (long double)M_LN2);
    if (out[16] == 0.0) out[16] = 0.0; /* canonical +0 */
}

/* floor(255*x/vmax) == ((x/vmax)*255 -> uint8) for every uint16 pair */
static inline uint8_t quantise1(uint16_t x, uint16_t vmax) {
    if (vmax == 0) return 0;
    return (uint8_t)(((uint32_t)x * 255u) / vmax);
}

static void glcm_offset(int angle_idx, int *dr, int *dc) {
    /* d=5: angles 0, pi/4, pi/2, 3pi/4 -> (0,5) (4,4) (5,0) (4,-4)   */
    static const int DR[4] = {0, 4, 5, 4};
    static const int DC[4] = {5, 4, 0, -4};
    *dr = DR[angle_idx];
    *dc = DC[angle_idx];
}

/* GLCM counts of one plane / one direction into P[256*256] (uint32), returns #pairs. */
static long glcm_count(const uint8_t *q, const uint8_t *mask, int h, int w, int angle_idx,
                       uint32_t *P) {
    int dr, dc, r, c;
    long M = 0;
    glcm_offset(angle_idx, &dr, &dc);
    memset(P, 0, sizeof(uint32_t) * LEVELS * LEVELS);
    for (r = 0; r < h; ++r) {
        int r2 = r + dr;
        if (r2 < 0 || r2 >= h) continue;
        for (c = 0; c < w; ++c) {
            int c2 = c + dc;
            if (c2 < 0 || c2 >= w) continue;
            if (mask && !(mask[r * w + c] && mask[r2 * w + c2])) continue;
            P[(int)q[r * w + c] * LEVELS + q[r2 * w + c2]]++;
            ++M;
        }
    }
    return M;
}

static void glcm_props(const uint32_t *P, long M, double *out) {
    double s = (M == 0) ? 1.0 : (double)M;
    long double con = 0, dis = 0, hom = 0, asm_ = 0, mui = 0, muj = 0;
    int i, j;
    for (i = 0; i < LEVELS; ++i)
        for (j = 0; j < LEVELS; ++j) {
            uint32_t cnt = P[i * LEVELS + j];
            if (!cnt) continue;
            double p = (double)cnt / s;
            double d = (double)(i - j);
            con += p * d * d;
            dis += p * fabs(d);
            hom += p / (1.0 + d * d);
            asm_ += p * p;
            mui += (double)i * p;
            muj += (double)j * p;
        }
    long double vi = 0, vj = 0, cov = 0;
    for (i = 0; i < LEVELS; ++i)
        for (j = 0; j < LEVELS; ++j) {
            uint32_t cnt = P[i * LEVELS + j];
            if (!cnt) continue;
            double p = (double)cnt / s;
            double di = (double)i - (double)mui, dj = (double)j - (double)muj;
            vi += p * di * di;
            vj += p * dj * dj;
            cov += p * di * dj;
        }
    double si = sqrt((double)vi), sj = sqrt((double)vj);
    out[0] = (double)con;
    out[1] = (double)dis;
    out[2] = (double)hom;
    out[3] = (double)asm_;
    out[4] = sqrt((double)asm_);
    out[5] = (si < 1e-15 || sj < 1e-15) ? 1.0 : (double)cov / (si * sj);
}

static void shape_block(const uint8_t *mask, int h, int w, double *out) {
    /* mask == NULL means "all pixels" */
    long area = 0;
    int64_t sr = 0, sc = 0, srr = 0, scc = 0, src = 0;
    int rmin = h, rmax = -1, cmin = w, cmax = -1, r, c, k;
    long n1 = 0, n2 = 0, n3 = 0;
#define MK(rr, cc) (((rr) < 0 || (rr) >= h || (cc) < 0 || (cc) >= w) ? 0 : (mask ? (mask[(rr)*w + (cc)] != 0) : 1))
#define BD(rr, cc) (MK(rr, cc) && !(MK((rr)-1, cc) && MK((rr) + 1, cc) && MK(rr, (cc)-1) && MK(rr, (cc) + 1)))
    for (r = 0; r < h; ++r)
        for (c = 0; c < w; ++c) {
            if (!MK(r, c)) continue;
            ++area;
            sr += r; sc += c; srr += (int64_t)r * r; scc += (int64_t)c * c; src += (int64_t)r * c;
            if (r < rmin) rmin = r;
            if (r > rmax) rmax = r;
            if (c < cmin) cmin = c;
            if (c > cmax) cmax = c;
        }
    /* skimage.measure.perimeter(neighbourhood=4): classify EVERY pixel by the weighted
       3x3 sum of the border image */
    for (r = 0; r < h; ++r)
        for (c = 0; c < w; ++c) {
            int v = 10 * (BD(r - 1, c - 1) + BD(r - 1, c + 1) + BD(r + 1, c - 1) + BD(r + 1, c + 1)) +
                    2 * (BD(r - 1, c) + BD(r + 1, c) + BD(r, c - 1) + BD(r, c + 1)) + BD(r, c);
            if (v == 5 || v == 7 || v == 15 || v == 17 || v == 25 || v == 27) ++n1;
            else if (v == 21 || v == 33) ++n2;
            else if (v == 13 || v == 23) ++n3;
        }
#undef BD
#undef MK
    double perim = (double)n1 + (double)n2 * sqrt(2.0) + (double)n3 * ((1.0 + sqrt(2.0)) / 2.0);
    out[0] = (double)area;
    out[1] = perim;
    if (area == 0) {
        for (k = 2; k < N_SHAPE; ++k) out[k] = NAN;
        return;
    }
    double A = (double)area, A2 = A * A;
    double bbox = (double)(rmax - rmin + 1) * (double)(cmax - cmin + 1);
    int64_t an = (int64_t)area * scc - sc * sc, cn = (int64_t)area * srr - sr * sr,
            bn = (int64_t)area * src - sr * sc;
    double a = (double)an / A2, cc_ = (double)cn / A2, b = -(double)bn / A2;
    double hd = (double)(an - cn) / A2 * 0.5;
    double D = sqrt(hd * hd + b * b);
    double l1 = (a + cc_) * 0.5 + D, l2 = 0.0, ecc = 0.0;
    if (l1 > 0) {
        l2 = (a * cc_ - b * b) / l1;
        if (l2 < 0) l2 = 0;
        ecc = 2.0 * D / l1;
        if (ecc < 0) ecc = 0;
        if (ecc > 1) ecc = 1;
        ecc = sqrt(ecc);
    }
    out[2] = bbox;
    out[3] = A / bbox;
    out[4] = (double)sr / A;
    out[5] = (double)sc / A;
    out[6] = 4.0 * sqrt(l1);
    out[7] = 4.0 * sqrt(l2);
    out[8] = ecc;
    out[9] = perim > 0 ? 4.0 * M_PI * A / (perim * perim) : NAN;
}

static void moment_block(const uint16_t *px, const uint8_t *mask, int h, int w, double *out) {
    long double m00 = 0, m10 = 0, m01 = 0;
    int r, c, k;
    for (r = 0; r < h; ++r)
        for (c = 0; c < w; ++c) {
            if (mask && !mask[r * w + c]) continue;
            double v = px[r * w + c];
            m00 += v; m10 += v * r; m01 += v * c;
        }
    if (m00 == 0) {
        for (k = 0; k < N_MOM; ++k) out[k] = NAN;
        return;
    }
    double cr = (double)(m10 / m00), cc = (double)(m01 / m00);
    long double mu[4][4];
    memset(mu, 0, sizeof(mu));
    for (r = 0; r < h; ++r)
        for (c = 0; c < w; ++c) {
            if (mask && !mask[r * w + c]) continue;
            double v = px[r * w + c];
            double dr = r - cr, dc = c - cc;
            mu[2][0] += dr * dr * v; mu[1][1] += dr * dc * v; mu[0][2] += dc * dc * v;
            mu[3][0] += dr * dr * dr * v; mu[2][1] += dr * dr * dc * v;
            mu[1][2] += dr * dc * dc * v; mu[0][3] += dc * dc * dc * v;
        }
    double M = (double)m00, n2 = M * M, n3 = pow(M, 2.5);
    out[0] = cr; out[1] = cc;
    out[2] = (double)mu[2][0] / n2; out[3] = (double)mu[1][1] / n2; out[4] = (double)mu[0][2] / n2;
    out[5] = (double)mu[3][0] / n3; out[6] = (double)mu[2][1] / n3;
    out[7] = (double)mu[1][2] / n3; out[8] = (double)mu[0][3] / n3;
}

/* Features of one plane.  mask may be NULL.  out_* may be NULL to skip a block.
   counts (optional): uint32[n_angles][256*256] raw GLCM bins. */
void imfeat_ref_plane(const uint16_t *px, const uint8_t *mask, int h, int w, int n_angles,
                      double *out_basic, double *out_glcm, double *out_shape, double *out_mom,
                      uint32_t *counts) {
    long n = (long)h * w, i, nsel = 0;
    uint16_t *sel = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)(n > 0 ? n : 1));
    uint16_t vmax = 0;
    for (i = 0; i < n; ++i)
        if (!mask || mask[i]) {
            sel[nsel++] = px[i];
            if (px[i] > vmax) vmax = px[i];
        }
    if (out_basic) basic_block(sel, nsel, out_basic);
    free(sel);
    if (out_glcm || counts) {
        uint8_t *q = (uint8_t *)malloc((size_t)(n > 0 ? n : 1));
        uint32_t *P = counts ? NULL : (uint32_t *)malloc(sizeof(uint32_t) * LEVELS * LEVELS);
        int a;
        for (i = 0; i < n; ++i) q[i] = (!mask || mask[i]) ? quantise1(px[i], vmax) : 0;
        for (a = 0; a < n_angles; ++a) {
            uint32_t *Pa = counts ? counts + (size_t)a * LEVELS * LEVELS : P;
            long M = glcm_count(q, mask, h, w, a, Pa);
            if (out_glcm) glcm_props(Pa, M, out_glcm + a * N_GLCM);
        }
        free(q);
        if (P) free(P);
    }
    if (out_shape) shape_block(mask, h, w, out_shape);
    if (out_mom) moment_block(px, mask, h, w, out_mom);
}

/* Whole table.  planes: plane-compact planar layout uint16[N][C][plane_stride] with the
   valid h_i x w_i region stored row-major at the start of each plane (the layout the
   product's batcher produces); masks same layout uint8 or NULL; sizes int32[N][2] or NULL
   (then every object is Hs x Ws).  out: double[N][F_total] in notebook column order:
   basic block (17*C), GLCM block (6*n_angles*C), shape block (10*C), moment block (9*C). */
void imfeat_ref_table(const uint16_t *planes, const uint8_t *masks, const int32_t *sizes,
                      long N, int C, int Hs, int Ws, long plane_stride, int want_glcm,
                      int n_angles, int want_shape, int want_mom, double *out) {
    int na = want_glcm ? n_angles : 0;
    long F = (long)C * (N_BASIC + N_GLCM * na + (want_shape ? N_SHAPE : 0) + (want_mom ? N_MOM : 0));
    long obj;
    for (obj = 0; obj < N; ++obj) {
        int h = sizes ? sizes[2 * obj] : Hs, w = sizes ? sizes[2 * obj + 1] : Ws, ch;
        double *row = out + obj * F;
        for (ch = 0; ch < C; ++ch) {
            const uint16_t *px = planes + ((size_t)obj * C + ch) * (size_t)plane_stride;
            const uint8_t *mk = masks ? masks + ((size_t)obj * C + ch) * (size_t)plane_stride : NULL;
            long off = (long)C * N_BASIC;
            double *ob = row + (long)ch * N_BASIC;
            double *og = na ? row + off + (long)ch * N_GLCM * na : NULL;
            off += (long)C * N_GLCM * na;
            double *os = want_shape ? row + off + (long)ch * N_SHAPE : NULL;
            off += want_shape ? (long)C * N_SHAPE : 0;
            double *om = want_mom ? row + off + (long)ch * N_MOM : NULL;
            imfeat_ref_plane(px, mk, h, w, na ? na : 1, ob, og, os, om, NULL);
        }
    }
}
