/*
 * imfeat.h -- C ABI of the B200-native per-object, per-channel feature extraction.
 *
 * This is the drop-in boundary for ONE path of
 * aliechoes/interpretable-multichannel-image-analysis: the body of the extraction loop of
 * channel_importance_hand_crafted_features.ipynb.  "NB:<n>" = raw JSON line n of that file.
 *
 *   reference interface                                   replaced by
 *   ---------------------------------------------------   ------------------------------
 *   basic_statistical_features(image)       NB:220-264    imfeat_extract_* (want_basic)
 *   glcm_features(image)                    NB:269-308    imfeat_extract_* (want_glcm)
 *   for i in tqdm(...): ... df.loc[i,:]=f   NB:358-364    one imfeat_extract_* call per batch
 *   greycomatrix(...) bins                  NB:298        imfeat_glcm_counts_device (parity/debug)
 *   README.md:8-9  image (h,w,c) 16-bit, mask (h,w,c)     imfeat_pack_hwc_* (layout conversion)
 *
 * The reference is pure Python, so a maintainer binds this with ctypes (INTEGRATION.md shows
 * the stub).  All entry points take plain pointers and sizes; no torch / C++ types.
 *
 * Data layout ("plane-compact planar"): planes is uint16[N][C][plane_stride]; plane (i, c)
 * holds the valid h_i x w_i region of channel c of object i ROW-MAJOR AND COMPACT (row pitch
 * w_i) in its first h_i*w_i elements; the rest of the plane is padding and is never read.
 * plane_stride is a multiple of 8 elements and planes/masks are 16-byte aligned.  sizes is
 * int32[N][2] = (h_i, w_i), or NULL when every object is Hs x Ws.  masks, when given, is
 * uint8 in the same layout (non-zero = inside).  h_i*w_i <= IMFEAT_MAX_PIXELS.
 *
 * Output: float64[N][row_stride]; the first imfeat_row_width() entries of a row are, in the
 * notebook's column order (NB:330-331):
 *   [0, 17*C)                        basic block, 17 per channel slot  (NB:241-262)
 *   [.., + 6*A*C)                    GLCM block, A angle groups of 6 per slot (NB:301-306)
 *   [.., + IMFEAT_N_SHAPE*C)         shape block        (extension, no reference counterpart)
 *   [.., + IMFEAT_N_MOMENT*C)        spatial moments    (extension, no reference counterpart)
 * Integer-valued features (min, max, total, area, ...) are exact integers stored as doubles.
 *
 * Every function returns IMFEAT_OK or a negative error code; imfeat_last_error() gives text.
 * Data-dependent degeneracy is never an error: the table gets the NaN / special values the
 * reference's CPU path produces and a bit is set in status[i].
 */
#ifndef IMFEAT_H
#define IMFEAT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IMFEAT_ABI_VERSION 1

#define IMFEAT_OK 0
#define IMFEAT_ERR_ARG (-1)       /* bad shape / null pointer / unsupported option */
#define IMFEAT_ERR_CUDA (-2)      /* a CUDA call failed; see imfeat_last_error */
#define IMFEAT_ERR_NOMEM (-3)

#define IMFEAT_N_BASIC 17
#define IMFEAT_N_GLCM 6
#define IMFEAT_N_SHAPE 10
#define IMFEAT_N_MOMENT 9
#define IMFEAT_MAX_ANGLES 4
#define IMFEAT_MAX_PIXELS 32768   /* per plane: 16-bit counters in the shared-memory tables */
#define IMFEAT_GLCM_LEVELS 256

/* status bits, per output row */
#define IMFEAT_ST_EMPTY_MASK 1u   /* some channel had no pixel inside its mask (NaN block)  */
#define IMFEAT_ST_NO_PAIRS 2u     /* some GLCM direction had zero pixel pairs               */
#define IMFEAT_ST_CONSTANT 4u     /* some channel is constant (NaN skew/kurtosis, NB:259)   */

typedef struct imfeat_ctx imfeat_ctx;

typedef struct imfeat_opts {
    int32_t struct_size;          /* = sizeof(imfeat_opts) */
    int32_t want_basic;           /* 17 basic columns per slot (NB:241-262)                 */
    int32_t want_glcm;            /* 6 GLCM columns per angle per slot (NB:293-306)         */
    int32_t n_angles;             /* 1 (notebook: angle 0) .. 4 (0, 45, 90, 135 degrees)    */
    int32_t glcm_distance;        /* notebook literal: 5 (NB:298)                           */
    int32_t want_shape;           /* extension                                              */
    int32_t want_moments;         /* extension                                              */
    int32_t host_mask_bits;       /* host entry points only: 1 = h_masks is bit-packed (see below)      */
    double percentiles[9];        /* np.percentile q arguments; notebook: 0.1 .. 0.9        */
} imfeat_opts;

/* Fills the notebook's literals: basic + GLCM(d=5, angle 0), q = 0.1..0.9, no extensions. */
void imfeat_default_opts(imfeat_opts *opts);

int imfeat_abi_version(void);

/* Context: owns per-device lookup tables and staging buffers.  Not thread-safe per ctx. */
int imfeat_create(int device, imfeat_ctx **out_ctx);
int imfeat_destroy(imfeat_ctx *ctx);
const char *imfeat_last_error(const imfeat_ctx *ctx);   /* ctx may be NULL */

/* Number of valid doubles per output row for C_out channel slots. */
int64_t imfeat_row_width(int32_t c_out, const imfeat_opts *opts);

/*
 * The hot path on device-resident data (replaces NB:358-364 for one batch).
 *   d_planes/d_masks/d_sizes : inputs as described above (d_masks, d_sizes may be NULL)
 *   d_src_obj : int32[N][c_out] or NULL. Row i, slot j reads object d_src_obj[i][j]
 *               (channel permutation across objects for importance sweeps); NULL = i.
 *   d_chan    : int32[c_out] or NULL. Slot j reads physical channel d_chan[j]
 *               (leave-one-channel-out = a shorter list); NULL = identity, c_out == c_in.
 *   d_out     : float64[N][row_stride], row_stride >= imfeat_row_width(c_out, opts)
 *   d_status  : uint32[N] or NULL
 *   stream    : cudaStream_t (0 = default stream). The call only enqueues work.
 */
int imfeat_extract_device(imfeat_ctx *ctx, const uint16_t *d_planes, const uint8_t *d_masks,
                          const int32_t *d_sizes, const int32_t *d_src_obj,
                          const int32_t *d_chan, int64_t n_objects, int32_t c_in,
                          int32_t c_out, int32_t hs, int32_t ws, int64_t plane_stride,
                          const imfeat_opts *opts, double *d_out, int64_t row_stride,
                          uint32_t *d_status, void *stream);

/*
 * Same computation from HOST buffers in the same plane-compact planar layout: stages through
 * pinned memory, copies host->device in slabs overlapped with the kernels, and copies the
 * table (and status) back.  Synchronous.  h_masks / h_sizes / h_status may be NULL.
 */
int imfeat_extract_host(imfeat_ctx *ctx, const uint16_t *h_planes, const uint8_t *h_masks,
                        const int32_t *h_sizes, int64_t n_objects, int32_t c, int32_t hs,
                        int32_t ws, int64_t plane_stride, const imfeat_opts *opts,
                        double *h_out, int64_t row_stride, uint32_t *h_status);

/*
 * Same, from the reference's own object layout (README.md:8-9): interleaved uint16[N][hs][ws][c]
 * images and uint8 masks of the same shape; object i occupies the top-left h_i x w_i corner
 * when h_sizes is given.  The interleaved slab is copied to the device as is and converted to
 * the planar layout there (imfeat_pack_hwc_device), so the host never transposes pixels.
 * This is the call that replaces the whole loop body NB:358-364 for a batch of objects.
 *
 * Bit-packed host masks (opts->host_mask_bits = 1, both host entry points): the masks are a third of the bytes
 * that cross PCIe; packed they are 4%.  Element k of an object (imfeat_extract_host_hwc: k = (r*ws + col)*c + ch
 * over the padded hs x ws x c block) or of a plane (imfeat_extract_host: k = index inside the plane's
 * plane_stride elements) is bit k & 7 of byte k >> 3 (numpy.packbits(..., bitorder="little")); every object /
 * plane starts at a multiple of 8 bytes: IMFEAT_MASK_BITS_BYTES(n_elements) bytes each.  The masks are
 * expanded to bytes on the device.
 */
#define IMFEAT_MASK_BITS_BYTES(n_elements) ((((int64_t)(n_elements) + 63) / 64) * 8)
int imfeat_extract_host_hwc(imfeat_ctx *ctx, const uint16_t *h_hwc, const uint8_t *h_mask_hwc,
                            const int32_t *h_sizes, int64_t n_objects, int32_t c, int32_t hs,
                            int32_t ws, const imfeat_opts *opts, double *h_out,
                            int64_t row_stride, uint32_t *h_status);

/*
 * Parity/debug: the raw GLCM bins of NB:298 for every (row, slot, angle):
 * d_counts is uint32[N][c_out][n_angles][256][256].
 */
int imfeat_glcm_counts_device(imfeat_ctx *ctx, const uint16_t *d_planes, const uint8_t *d_masks,
                              const int32_t *d_sizes, int64_t n_objects, int32_t c,
                              int32_t hs, int32_t ws, int64_t plane_stride,
                              const imfeat_opts *opts, uint32_t *d_counts, void *stream);

/*
 * Layout conversion on the device: README.md:8-9 objects stored interleaved
 * uint16[N][hs][ws][c] (and uint8 masks) -> plane-compact planar uint16[N][c][plane_stride].
 * d_sizes (or NULL) gives the valid (h_i, w_i) top-left region of each padded object.
 */
int imfeat_pack_hwc_device(imfeat_ctx *ctx, const uint16_t *d_hwc, const uint8_t *d_mask_hwc,
                           const int32_t *d_sizes, int64_t n_objects, int32_t c, int32_t hs,
                           int32_t ws, int64_t plane_stride, uint16_t *d_planes,
                           uint8_t *d_masks, void *stream);

/*
 * Synthetic objects generated on the device with a counter-based hash (bit-identical numpy
 * mirror: <package>/synth.py).  Writes planes, masks (may be NULL) and, when variable != 0,
 * sizes drawn in [hmin..hs] x [wmin..ws]; otherwise every object is hs x ws and d_sizes may
 * be NULL.  first_object offsets the object counter (sharding).
 */
int imfeat_synth_device(imfeat_ctx *ctx, uint64_t seed, int64_t first_object,
                        int64_t n_objects, int32_t c, int32_t hs, int32_t ws,
                        int64_t plane_stride, int32_t variable, int32_t hmin, int32_t wmin,
                        int32_t mask_shrink_256, uint16_t *d_planes, uint8_t *d_masks,
                        int32_t *d_sizes, void *stream);

/*
 * Table post-processing on the device: the notebook's MinMaxScaler step
 * (channel_importance_hand_crafted_features.ipynb cell 16, NB:389-394:
 * `norm = MinMaxScaler().fit(X_train); X_train = norm.transform(X_train); X_test = norm.transform(X_test)`),
 * sklearn semantics (nanmin / nanmax per column, ranges below 10 eps scale by 1, X * scale_ + min_).
 *
 * imfeat_minmax_fit_device: d_table float64[n_rows][row_stride] (first n_cols cells of a row are used);
 * d_stats float64[4][n_cols] receives data_min_, data_max_, scale_, min_.
 * imfeat_minmax_transform_device: d_out[r][c] = d_in[r][c] * scale_[c] + min_[c]; d_out may alias d_in.
 */
int imfeat_minmax_fit_device(imfeat_ctx *ctx, const double *d_table, int64_t n_rows, int32_t n_cols,
                             int64_t row_stride, double *d_stats, void *stream);
int imfeat_minmax_transform_device(imfeat_ctx *ctx, const double *d_in, int64_t n_rows,
                                   int32_t n_cols, int64_t row_stride_in, const double *d_stats,
                                   double *d_out, int64_t row_stride_out, void *stream);

/*
 * Optional per-kernel device timing (CUDA events on the launching stream, resolved lazily, the
 * stream is not serialised).  imfeat_kernel_times returns, per kernel group
 * [0] K1 moments, [1] K2 order statistics + entropy, [2] K3 GLCM, [3] K4 shape/moments,
 * the accumulated milliseconds and launch counts since the last reset.
 */
int imfeat_enable_timing(imfeat_ctx *ctx, int32_t enable);
int imfeat_kernel_times(imfeat_ctx *ctx, double *ms_out, int64_t *calls_out, int32_t reset);

/* Number of kernel launches this context has enqueued so far (bench accounting). */
int64_t imfeat_launch_count(const imfeat_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* IMFEAT_H */
