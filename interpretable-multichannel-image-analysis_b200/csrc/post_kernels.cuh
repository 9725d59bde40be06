// post_kernels.cuh -- feature-table post-processing that the notebook does right after extraction:
//   norm = MinMaxScaler().fit(X_train); X_train = norm.transform(X_train); X_test = norm.transform(X_test)
//   (channel_importance_hand_crafted_features.ipynb cell 16, NB:389-394).
// Kept on the device so a table that stays there for repeated ablations never visits the host.
//
// sklearn 1.9 semantics (sklearn/preprocessing/_data.py, MinMaxScaler.partial_fit / transform, default
// feature_range (0, 1)): data_min = nanmin, data_max = nanmax per column; range = max - min;
// scale = 1 / range, with range < 10 * eps treated as 1; min_ = 0 - data_min * scale;
// transform: X * scale + min_ (two roundings).  NaN cells stay NaN; an all-NaN column gives NaN.
// HBM-bound: the table is row-major float64, threads run along columns (coalesced), rows are split over
// gridDim.y with one partial min/max per row block.
#pragma once
#include "common.cuh"

namespace imfeat {

constexpr int kPostRowBlocks = 64;

// partial[by][0][col] = min, partial[by][1][col] = max over rows by, by + gridDim.y, ... (NaN ignored)
__global__ void __launch_bounds__(128) colstats_partial_kernel(const double* __restrict__ table, long long n, int f,
                                                               long long row_stride, double* __restrict__ partial) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= f) return;
    double mn = __longlong_as_double(0x7ff0000000000000ll), mx = -mn;      // +inf, -inf
    for (long long r = blockIdx.y; r < n; r += gridDim.y) {
        const double v = table[r * row_stride + col];
        if (v == v) { mn = fmin(mn, v); mx = fmax(mx, v); }
    }
    partial[((long long)blockIdx.y * 2 + 0) * f + col] = mn;
    partial[((long long)blockIdx.y * 2 + 1) * f + col] = mx;
}

// out[0][col] = data_min, out[1][col] = data_max, out[2][col] = scale_, out[3][col] = min_
__global__ void __launch_bounds__(128) minmax_finish_kernel(const double* __restrict__ partial, int blocks, int f,
                                                            double* __restrict__ out) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= f) return;
    double mn = __longlong_as_double(0x7ff0000000000000ll), mx = -mn;
    for (int b = 0; b < blocks; ++b) {
        mn = fmin(mn, partial[((long long)b * 2 + 0) * f + col]);
        mx = fmax(mx, partial[((long long)b * 2 + 1) * f + col]);
    }
    if (mn > mx) { mn = qnan(); mx = qnan(); }               // no finite-or-infinite value at all: nanmin gives NaN
    double range = __dsub_rn(mx, mn);
    if (range < 10.0 * 2.220446049250313e-16) range = 1.0;   // _handle_zeros_in_scale
    const double scale = __ddiv_rn(1.0, range);
    out[col] = mn;
    out[f + col] = mx;
    out[2 * f + col] = scale;
    out[3 * f + col] = __dsub_rn(0.0, __dmul_rn(mn, scale));
}

__global__ void __launch_bounds__(256) minmax_apply_kernel(const double* __restrict__ in, long long n, int f,
                                                           long long stride_in, const double* __restrict__ stats,
                                                           double* __restrict__ out, long long stride_out) {
    const long long total = n * f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / f;
        const int col = (int)(i - r * f);
        const double v = in[r * stride_in + col];
        out[r * stride_out + col] = __dadd_rn(__dmul_rn(v, stats[2 * f + col]), stats[3 * f + col]);
    }
}

}  // namespace imfeat
