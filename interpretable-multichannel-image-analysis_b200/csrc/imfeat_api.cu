// imfeat_api.cu -- the C ABI declared in include/imfeat.h (single translation unit, sm_100a).
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/imfeat.h"
#include "aux_kernels.cuh"
#include "post_kernels.cuh"
#include "common.cuh"
#include "k1_moments.cuh"
#include "k2_order_entropy.cuh"
#include "k12_basic.cuh"
#include "k2b_bytes.cuh"
#include "k3_glcm.cuh"
#include "k3_ring.cuh"
#include "k4_shape.cuh"

using namespace imfeat;

static_assert(kNBasic == IMFEAT_N_BASIC && kNGlcm == IMFEAT_N_GLCM, "header mismatch");
static_assert(kNShape == IMFEAT_N_SHAPE && kNMoment == IMFEAT_N_MOMENT, "header mismatch");
static_assert(kMaxPixels == IMFEAT_MAX_PIXELS && kMaxAngles == IMFEAT_MAX_ANGLES, "header mismatch");

constexpr int kTimingSlots = 64;
constexpr int kSchedSlots = 256;
constexpr int kWlSlots = 8;

struct imfeat_ctx {
    int device;
    int sm_count;
    int k1_bps[2], k4_bps[2], k2c_bps[2], k4w_bps[2], k12_bps[2];   // resident CTAs per SM (occupancy API), [masked]
    unsigned int* d_sched;      // ring of kSchedSlots x 8 work counters (one slot per extract call)
    unsigned int sched_head;
    uint32_t* d_worklist;       // [0] = count, [1..] = tile ids left to the full-range K2 kernel
    size_t worklist_cap;        // entries per ring slot
    unsigned int wl_head;
    void* retired[64];          // outgrown work buffers: kept until imfeat_destroy (a captured graph or a call in
    int n_retired;              // flight on another stream may still hold their addresses)
    // K3 scratch (quantised tiles between the front and the bins kernel): two slots, handed out in turn; a slot
    // is reused only after the event recorded behind its last consumer
    struct { unsigned char* ptr; size_t bytes; cudaEvent_t ev; int recorded; } scr[2];
    unsigned int scr_head;
    double* d_log2tab;
    unsigned long long* d_gfix;
    long long launches;
    char err[512];
    // host-path staging (lazily sized)
    cudaStream_t streams[2];
    cudaEvent_t done[2];
    void* pin_in[2];
    void* pin_out[2];
    void* dev_in[2];
    void* dev_out[2];
    size_t in_bytes, out_bytes;
    // optional per-kernel timing (imfeat_enable_timing): a ring of event sets, resolved lazily
    // debugging / measurement switches, read from the environment once at imfeat_create
    int env_k1_fp64, env_k1_tma, env_k2_compact, env_k4_warp, env_k2_groups, env_k3_threads, env_k3_chunk, env_k3_ring, env_k3_table, env_fuse12;
    int timing;
    int t_head, t_pending;
    cudaEvent_t t_ev[kTimingSlots][5];
    cudaEvent_t t_side[kTimingSlots][2];   // K4 on the side stream (overlap mode): its own start / end
    // overlap mode: K4w runs on a side stream next to K3 (its few resident warps fill the issue slots the
    // shared-memory-bound GLCM kernels leave free)
    int env_overlap, env_k4_fill, env_k3_tiers, env_k2_bytes;
    int k2b_bps[2];
    cudaStream_t side[4];
    cudaEvent_t fork_ev[8], join_ev[8];
    unsigned int side_head;
    unsigned t_mask[kTimingSlots];
    double t_ms[4];
    long long t_calls[4];
};

static thread_local char g_err[512] = "";

static int fail(imfeat_ctx* ctx, int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    if (ctx) memcpy(ctx->err, g_err, sizeof(g_err));
    return code;
}

// Every entry point works on the context's device and puts the caller's current device back on return
// (PyTorch reads the current device through cudaGetDevice: leaving it changed would silently move the
// calling thread's later allocations, streams and NCCL calls to another GPU).
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

#define CU(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return fail(ctx, IMFEAT_ERR_CUDA, "%s failed: %s (%s:%d)", #call,             \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                      \
    } while (0)

extern "C" {

void imfeat_default_opts(imfeat_opts* o) {
    memset(o, 0, sizeof(*o));
    o->struct_size = (int32_t)sizeof(*o);
    o->want_basic = 1;
    o->want_glcm = 1;
    o->n_angles = 1;
    o->glcm_distance = 5;
    for (int k = 0; k < 9; ++k) o->percentiles[k] = (double)(k + 1) / 10.0;
}

int imfeat_abi_version(void) { return IMFEAT_ABI_VERSION; }

const char* imfeat_last_error(const imfeat_ctx* ctx) { return ctx ? ctx->err : g_err; }

int64_t imfeat_launch_count(const imfeat_ctx* ctx) { return ctx ? ctx->launches : 0; }

int64_t imfeat_row_width(int32_t c_out, const imfeat_opts* o) {
    if (!o || c_out < 0) return -1;
    int64_t per = 0;
    if (o->want_basic) per += kNBasic;
    if (o->want_glcm) per += kNGlcm * o->n_angles;
    if (o->want_shape) per += kNShape;
    if (o->want_moments) per += kNMoment;
    return per * c_out;
}

int imfeat_create(int device, imfeat_ctx** out_ctx) {
    imfeat_ctx* ctx = nullptr;
    if (!out_ctx) return fail(nullptr, IMFEAT_ERR_ARG, "out_ctx is NULL");
    *out_ctx = nullptr;
    int count = 0;
    CU(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count)
        return fail(nullptr, IMFEAT_ERR_ARG, "device %d out of range (%d CUDA devices)", device, count);
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(nullptr, IMFEAT_ERR_ARG,
                    "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                    prop.major, prop.minor);
    ctx = (imfeat_ctx*)calloc(1, sizeof(imfeat_ctx));
    if (!ctx) return fail(nullptr, IMFEAT_ERR_NOMEM, "out of host memory");
    ctx->device = device;
    {
        auto flag = [](const char* name, int def) { const char* e = getenv(name); return e ? atoi(e) : def; };
        ctx->env_k1_fp64 = flag("IMFEAT_K1_FP64", 0);      // 1: skip K1's integer pass (exercises the FP64 pass)
        ctx->env_k1_tma = flag("IMFEAT_K1_TMA", 0);        // 1: K1 through the cp.async.bulk ring (unmasked)
        ctx->env_k2_compact = flag("IMFEAT_K2_COMPACT", 1); // 0: every tile through the full-range ring kernel
        ctx->env_k4_warp = flag("IMFEAT_K4_WARP", 1);       // 0: CTA-per-tile K4 for every batch
        const int g2 = flag("IMFEAT_K2_GROUPS", 2);
        ctx->env_k2_groups = (g2 == 2 || g2 == 4 || g2 == 8) ? g2 : 2;
        ctx->env_k3_threads = flag("IMFEAT_K3_THREADS", 128) == 256 ? 256 : 128;   // threads per K3 CTA
        ctx->env_k3_chunk = flag("IMFEAT_K3_CHUNK", 16384);                        // objects per front/bins round of K3
        if (ctx->env_k3_chunk < 1) ctx->env_k3_chunk = 16384;
        ctx->env_fuse12 = flag("IMFEAT_FUSE12", 1);          // 1: the basic block in one pass (k12_basic.cuh); 0: K1 then K2c
        ctx->env_k3_table = flag("IMFEAT_K3_TABLE", 0);      // 32 / 64: force the table size of the bins kernel (0: by mask)
        ctx->env_k3_ring = flag("IMFEAT_K3_RING", 1);       // 1: unmasked tiles through the one-kernel ring variant (k3_ring.cuh)
        ctx->env_overlap = flag("IMFEAT_OVERLAP", 0);       // 1: K4w on a side stream next to K3, 2: next to K12 as well (measured: no gain); 0: one stream
        ctx->env_k4_fill = flag("IMFEAT_K4_FILL", 4);       // resident K4w warps per SM while it runs next to K3
        if (ctx->env_k4_fill < 1) ctx->env_k4_fill = 1;
        ctx->env_k2_bytes = flag("IMFEAT_K2_BYTES", 1);      // 0: full-range tiles straight to K2's single 16-bit table
        ctx->env_k3_tiers = flag("IMFEAT_K3_TIERS", 1);     // 0: K3 with room for a whole tile only (no first tier for large strides)
    }
    ctx->sm_count = prop.multiProcessorCount;
    // log2 table, computed on the host in double precision (k = 0 maps to 0, never used)
    double* tab = (double*)malloc(sizeof(double) * (kMaxPixels + 1));
    if (!tab) { free(ctx); return fail(nullptr, IMFEAT_ERR_NOMEM, "out of host memory"); }
    tab[0] = 0.0;
    for (int k = 1; k <= kMaxPixels; ++k) tab[k] = log2((double)k);
    cudaError_t e = cudaMalloc(&ctx->d_log2tab, sizeof(double) * (kMaxPixels + 1));
    if (e == cudaSuccess)
        e = cudaMemcpy(ctx->d_log2tab, tab, sizeof(double) * (kMaxPixels + 1), cudaMemcpyHostToDevice);
    // G[k] = (k+1)*log2(k+1) - k*log2(k) in 2^-42 fixed point (extended precision, rounded once);
    // G < 16.5, so a sum over kMaxPixels terms stays below 2^63
    unsigned long long* gfix = (unsigned long long*)tab;
    static_assert(sizeof(unsigned long long) == sizeof(double), "table reuse");
    for (int k = 0; k < kMaxPixels; ++k) {
        const long double a = (long double)(k + 1) * log2l((long double)(k + 1));
        const long double b = k ? (long double)k * log2l((long double)k) : 0.0L;
        gfix[k] = (unsigned long long)llroundl((a - b) * 4398046511104.0L);
    }
    if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->d_sched, sizeof(unsigned int) * 8 * kSchedSlots);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_gfix, sizeof(unsigned long long) * kMaxPixels);
    if (e == cudaSuccess)
        e = cudaMemcpy(ctx->d_gfix, gfix, sizeof(unsigned long long) * kMaxPixels, cudaMemcpyHostToDevice);
    free(tab);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k2_order_entropy_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K2Smem));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k2_order_entropy_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K2Smem));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k2b_order_entropy_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K2bSmem));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k2b_order_entropy_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K2bSmem));
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->k2b_bps[0], k2b_order_entropy_kernel<false>, kK2bThreads, sizeof(K2bSmem));
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->k2b_bps[1], k2b_order_entropy_kernel<true>, kK2bThreads, sizeof(K2bSmem));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k3_glcm_kernel<false, false, 128, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k3_glcm_kernel<false, false, 128, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k3_glcm_kernel<false, true, 128, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k3_glcm_kernel<false, true, 128, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k3_glcm_kernel<true, false, 128, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k3_glcm_kernel<true, false, 128, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k3_glcm_kernel<true, true, 128, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k3_glcm_kernel<true, true, 128, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k3_glcm_kernel<false, false, 256, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k3_glcm_kernel<false, false, 256, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k3_glcm_kernel<false, true, 256, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k3_glcm_kernel<false, true, 256, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k3_glcm_kernel<true, false, 256, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k3_glcm_kernel<true, false, 256, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k3_glcm_kernel<true, true, 256, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k3_glcm_kernel<true, true, 256, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ring::k3_glcm_kernel<false, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ring::k3_glcm_kernel<false, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k3a_front_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k3a_front_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k4_shape_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K4Smem));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k4_shape_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K4Smem));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k1_moments_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->k1_bps[0], k1_moments_kernel<false>, 256, 0);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->k1_bps[1], k1_moments_kernel<true>, 256, 0);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->k2c_bps[0], k2c_order_entropy_kernel<false>, kK2cThreads, 0);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->k2c_bps[1], k2c_order_entropy_kernel<true>, kK2cThreads, 0);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->k12_bps[0], k12_basic_kernel<false>, 32, 0);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->k12_bps[1], k12_basic_kernel<true>, 32, 0);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->k4w_bps[0], k4w_shape_kernel<false, true>, 32, 0);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->k4w_bps[1], k4w_shape_kernel<true, true>, 32, 0);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->k4_bps[0], k4_shape_kernel<false>, kK4Threads, sizeof(K4Smem));
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->k4_bps[1], k4_shape_kernel<true>, kK4Threads, sizeof(K4Smem));
    if (e != cudaSuccess) {
        int rc = fail(nullptr, IMFEAT_ERR_CUDA, "context set-up failed: %s", cudaGetErrorString(e));
        if (ctx->d_log2tab) cudaFree(ctx->d_log2tab);
        if (ctx->d_gfix) cudaFree(ctx->d_gfix);
        free(ctx);
        return rc;
    }
    *out_ctx = ctx;
    return IMFEAT_OK;
}

static void free_staging(imfeat_ctx* ctx) {
    for (int b = 0; b < 2; ++b) {
        if (ctx->pin_in[b]) cudaFreeHost(ctx->pin_in[b]);
        if (ctx->pin_out[b]) cudaFreeHost(ctx->pin_out[b]);
        if (ctx->dev_in[b]) cudaFree(ctx->dev_in[b]);
        if (ctx->dev_out[b]) cudaFree(ctx->dev_out[b]);
        ctx->pin_in[b] = ctx->pin_out[b] = ctx->dev_in[b] = ctx->dev_out[b] = nullptr;
    }
    ctx->in_bytes = ctx->out_bytes = 0;
}

int imfeat_destroy(imfeat_ctx* ctx) {
    if (!ctx) return IMFEAT_OK;
    DeviceGuard guard(ctx->device);
    free_staging(ctx);
    for (int b = 0; b < 2; ++b) {
        if (ctx->streams[b]) cudaStreamDestroy(ctx->streams[b]);
        if (ctx->done[b]) cudaEventDestroy(ctx->done[b]);
    }
    for (int sl = 0; sl < kTimingSlots; ++sl) {
        for (int k = 0; k < 5; ++k)
            if (ctx->t_ev[sl][k]) cudaEventDestroy(ctx->t_ev[sl][k]);
        for (int k = 0; k < 2; ++k)
            if (ctx->t_side[sl][k]) cudaEventDestroy(ctx->t_side[sl][k]);
    }
    for (int k = 0; k < 4; ++k) if (ctx->side[k]) cudaStreamDestroy(ctx->side[k]);
    for (int k = 0; k < 8; ++k) {
        if (ctx->fork_ev[k]) cudaEventDestroy(ctx->fork_ev[k]);
        if (ctx->join_ev[k]) cudaEventDestroy(ctx->join_ev[k]);
    }
    if (ctx->d_log2tab) cudaFree(ctx->d_log2tab);
    if (ctx->d_gfix) cudaFree(ctx->d_gfix);
    if (ctx->d_worklist) cudaFree(ctx->d_worklist);
    for (int k = 0; k < ctx->n_retired; ++k) cudaFree(ctx->retired[k]);
    for (int k = 0; k < 2; ++k) {
        if (ctx->scr[k].ptr) cudaFree(ctx->scr[k].ptr);
        if (ctx->scr[k].ev) cudaEventDestroy(ctx->scr[k].ev);
    }
    if (ctx->d_sched) cudaFree(ctx->d_sched);
    free(ctx);
    return IMFEAT_OK;
}

}  // extern "C"

// A work buffer that must grow is never freed or reused in place: a captured CUDA graph (ablation.CapturedSweep)
// or a call in flight on another stream may hold its address.  It is retired until imfeat_destroy; and nothing
// can be allocated while the stream is being captured.
static int grow_buffer(imfeat_ctx* ctx, cudaStream_t st, void** buf, size_t bytes, const char* what) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    CU(cudaStreamIsCapturing(st, &cap));
    if (cap != cudaStreamCaptureStatusNone)
        return fail(ctx, IMFEAT_ERR_ARG, "the batch outgrew the context's %s during stream capture; run one eager call "
                    "of this size first", what);
    if (*buf) {
        if (ctx->n_retired == 64) return fail(ctx, IMFEAT_ERR_NOMEM, "too many work-buffer generations");
        ctx->retired[ctx->n_retired++] = *buf;
        *buf = nullptr;
    }
    CU(cudaMalloc(buf, bytes));
    return IMFEAT_OK;
}

// K3 = front kernel (one warp per tile: quantised tile + geometry into the scratch record, pair sums into the
// output records), bins kernel (persistent CTAs of NT threads, as many as fit an SM: three for 64x64 tiles),
// finalize kernel (raw sums -> the six properties).  Front and bins alternate over chunks of objects so that the
// scratch records of a chunk are still in L2 when the bins kernel fetches them.
template <bool DUMP, int NT, int TB>
static int launch_k3_nt(imfeat_ctx* ctx, bool masked, cudaStream_t st, const Params& P, int maxpx) {
    // Capacity per tile.  Large strides with masks: two tiers (see K3Cap) -- the first holds the rows of a sparse
    // mask's bounding box and runs three times as many warps / CTAs per SM; the tiles it leaves over are listed
    // and taken by a second pair of launches with room for a whole tile.
    const int mb_px = ((P.hs * P.ws + 7) & ~7);
    const K3Cap cap_full = {maxpx, mb_px};
    int q1 = maxpx;
    if (masked && !DUMP && maxpx > 8192 && ctx->env_k3_tiers) {
        q1 = (maxpx / 3 + 15) & ~15;
        if (q1 < 4352) q1 = 4352;
    }
    const bool two = q1 < maxpx;
    const K3Cap cap1 = {q1, mb_px};
    struct Shape { size_t rec, smem_a, smem_b; int bps_a, bps_b; };
    auto shape_of = [&](K3Cap cap, Shape& sh) -> int {
        sh.rec = k3_rec_bytes(cap, masked);
        sh.smem_b = k3_smem_bytes(cap, masked, TB);
        sh.smem_a = (size_t)kK3aWarps * k3a_warp_bytes(cap, masked);
        sh.bps_a = sh.bps_b = 0;
        cudaError_t e = masked ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sh.bps_b, k3_glcm_kernel<true, DUMP, NT, TB>, NT, sh.smem_b)
                               : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sh.bps_b, k3_glcm_kernel<false, DUMP, NT, TB>, NT, sh.smem_b);
        if (e == cudaSuccess)
            e = masked ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sh.bps_a, k3a_front_kernel<true>, 32 * kK3aWarps, sh.smem_a)
                       : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sh.bps_a, k3a_front_kernel<false>, 32 * kK3aWarps, sh.smem_a);
        if (e != cudaSuccess) return fail(ctx, IMFEAT_ERR_CUDA, "K3 occupancy query failed: %s", cudaGetErrorString(e));
        if (sh.bps_a < 1 || sh.bps_b < 1)
            return fail(ctx, IMFEAT_ERR_CUDA, "K3 does not fit an SM (%zu / %zu bytes of shared memory)", sh.smem_a, sh.smem_b);
        return IMFEAT_OK;
    };
    Shape s1, s2;
    int rcs = shape_of(cap1, s1);
    if (rcs) return rcs;
    if (two) { rcs = shape_of(cap_full, s2); if (rcs) return rcs; }
    // chunks of whole objects, evenly sized
    const long long chunk_tiles = (long long)ctx->env_k3_chunk * P.c_out;
    const long long n_chunks = (P.n_tiles + chunk_tiles - 1) / chunk_tiles;
    const long long per = (P.n_tiles + n_chunks - 1) / n_chunks;
    // scratch slot: wait for its last consumer (possibly on another stream), grow if needed
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    CU(cudaStreamIsCapturing(st, &cap));
    auto& slot = ctx->scr[ctx->scr_head++ & 1u];
    if (cap == cudaStreamCaptureStatusNone && slot.recorded) CU(cudaStreamWaitEvent(st, slot.ev, 0));
    // layout: first-tier records, their lengths | the list of left-over tiles | second-tier records, their lengths
    const size_t off_wl = (((size_t)per * (s1.rec + sizeof(uint32_t))) + 255) & ~(size_t)255;
    const size_t off_r2 = two ? ((off_wl + (size_t)per * sizeof(uint32_t) + 255) & ~(size_t)255) : off_wl;
    const size_t scr_bytes = two ? off_r2 + (size_t)per * (s2.rec + sizeof(uint32_t)) : off_wl;
    if (slot.bytes < scr_bytes) {
        if (cap == cudaStreamCaptureStatusNone && slot.recorded) CU(cudaEventSynchronize(slot.ev));
        int rc = grow_buffer(ctx, st, (void**)&slot.ptr, scr_bytes, "K3 scratch");
        if (rc) { slot.bytes = 0; return rc; }
        slot.bytes = scr_bytes;
    }
    // work counters of one round (slots of the call's counter set): front / bins of tier 1, of tier 2, left-over count
    unsigned int* const c_a1 = P.sched + 5; unsigned int* const c_b1 = P.sched + 6;
    unsigned int* const c_a2 = P.sched + 2; unsigned int* const c_b2 = P.sched + 4; unsigned int* const c_left = P.sched + 7;
    auto launch_pair = [&](const Shape& sh, K3Cap cp, long long t0, uint32_t nl, unsigned char* scr, const K3Tier& ta, const K3Tier& tb) {
        const long long warps_a = (long long)ctx->sm_count * sh.bps_a * kK3aWarps;
        const int grid_a = (int)((nl < warps_a ? nl : warps_a) + kK3aWarps - 1) / kK3aWarps;
        const int pf = sh.bps_a * kK3aWarps < 16;           // few warps per SM (large strides): pull the next tile into L2
        if (masked) k3a_front_kernel<true><<<grid_a, 32 * kK3aWarps, sh.smem_a, st>>>(P, cp, (uint32_t)t0, nl, scr, pf, ta);
        else k3a_front_kernel<false><<<grid_a, 32 * kK3aWarps, sh.smem_a, st>>>(P, cp, (uint32_t)t0, nl, scr, pf, ta);
        const long long res_b = (long long)ctx->sm_count * sh.bps_b;
        const int grid_b = (int)(nl < res_b ? nl : res_b);
        if (masked) k3_glcm_kernel<true, DUMP, NT, TB><<<grid_b, NT, sh.smem_b, st>>>(P, cp, nl, scr, tb);
        else k3_glcm_kernel<false, DUMP, NT, TB><<<grid_b, NT, sh.smem_b, st>>>(P, cp, nl, scr, tb);
        ctx->launches += 2;
    };
    for (long long c = 0; c < n_chunks; ++c) {
        const long long t0 = c * per;
        const uint32_t nl = (uint32_t)(P.n_tiles - t0 < per ? P.n_tiles - t0 : per);
        CU(cudaMemsetAsync(P.sched + 5, 0, 2 * sizeof(unsigned int), st));     // the tile counters of the two kernels
        uint32_t* wl = reinterpret_cast<uint32_t*>(slot.ptr + off_wl);
        if (two) {
            CU(cudaMemsetAsync(c_a2, 0, sizeof(unsigned int), st));
            CU(cudaMemsetAsync(c_b2, 0, sizeof(unsigned int), st));
            CU(cudaMemsetAsync(c_left, 0, sizeof(unsigned int), st));
        }
        const K3Tier ta1 = {nullptr, nullptr, two ? wl : nullptr, two ? c_left : nullptr, c_a1};
        const K3Tier tb1 = {nullptr, nullptr, nullptr, nullptr, c_b1};
        launch_pair(s1, cap1, t0, nl, slot.ptr, ta1, tb1);
        if (two) {
            // the tiles the first tier listed (their number is known on the device only: the grids are sized for all)
            const K3Tier ta2 = {wl, c_left, nullptr, nullptr, c_a2};
            const K3Tier tb2 = {wl, c_left, nullptr, nullptr, c_b2};
            launch_pair(s2, cap_full, t0, nl, slot.ptr + off_r2, ta2, tb2);
        }
    }
    const long long recs = P.n_tiles * P.n_angles;
    const long long want = (recs + 255) / 256, capg = 8ll * ctx->sm_count;
    k3_finalize_kernel<<<(int)(want < capg ? want : capg), 256, 0, st>>>(P);
    ctx->launches += 1;
    if (cap == cudaStreamCaptureStatusNone) {
        if (!slot.ev) CU(cudaEventCreateWithFlags(&slot.ev, cudaEventDisableTiming));
        CU(cudaEventRecord(slot.ev, st));
        slot.recorded = 1;
    }
    return IMFEAT_OK;
}
template <bool DUMP>
static int launch_k3(imfeat_ctx* ctx, bool masked, cudaStream_t st, const Params& P, int maxpx) {
    // masked tiles: 32 KB tables of 4-bit counters (their bins rarely hold more than a few pairs); unmasked: 64 KB of 8-bit
    const bool small = ctx->env_k3_table ? ctx->env_k3_table == 32 : masked;
    if (ctx->env_k3_threads == 128)
        return small ? launch_k3_nt<DUMP, 128, 32>(ctx, masked, st, P, maxpx) : launch_k3_nt<DUMP, 128, 64>(ctx, masked, st, P, maxpx);
    return small ? launch_k3_nt<DUMP, 256, 32>(ctx, masked, st, P, maxpx) : launch_k3_nt<DUMP, 256, 64>(ctx, masked, st, P, maxpx);
}

// K2: measured on B200 (10,000 64x64x12 objects): unmasked 1.59 ms with 4 groups vs 1.84 ms with 2;
// masked (branch-free path) 2.19 ms vs 2.39 ms.
static int k2_groups(const imfeat_ctx* ctx) { return ctx->env_k2_groups; }

// C round(): half away from zero, as skimage's _glcm_loop uses for the pixel offsets.
static int c_round(double v) { return (int)(v < 0 ? -floor(-v + 0.5) : floor(v + 0.5)); }

static int check_common(imfeat_ctx* ctx, const void* planes, int64_t n, int c_in, int c_out, int hs,
                        int ws, int64_t plane_stride, const imfeat_opts* o, bool device_ptr = true) {
    if (!ctx) return fail(nullptr, IMFEAT_ERR_ARG, "ctx is NULL");
    if (!o || o->struct_size != (int32_t)sizeof(imfeat_opts))
        return fail(ctx, IMFEAT_ERR_ARG, "opts is NULL or has the wrong struct_size");
    if (n < 0) return fail(ctx, IMFEAT_ERR_ARG, "n_objects < 0");
    if (n > 0 && !planes) return fail(ctx, IMFEAT_ERR_ARG, "planes is NULL");
    if (c_in < 1 || c_out < 1) return fail(ctx, IMFEAT_ERR_ARG, "channel counts must be >= 1");
    if (n * (int64_t)c_out >= ((int64_t)1 << 31))
        return fail(ctx, IMFEAT_ERR_ARG, "n_objects * channels must be < 2^31 per call; split the batch");
    if (hs < 1 || ws < 1 || (int64_t)hs * ws > kMaxPixels)
        return fail(ctx, IMFEAT_ERR_ARG, "object size %dx%d outside 1..%d pixels per plane", hs, ws, kMaxPixels);
    if (plane_stride < (int64_t)hs * ws || (plane_stride & 7))
        return fail(ctx, IMFEAT_ERR_ARG, "plane_stride must be >= hs*ws and a multiple of 8");
    // only device planes are read with 128-bit loads; host buffers are copied (memcpy / DMA) first
    if (device_ptr && ((uintptr_t)planes & 15) != 0) return fail(ctx, IMFEAT_ERR_ARG, "planes must be 16-byte aligned");
    if (o->want_glcm && (o->n_angles < 1 || o->n_angles > kMaxAngles))
        return fail(ctx, IMFEAT_ERR_ARG, "n_angles must be 1..%d", kMaxAngles);
    if (o->want_glcm && (o->glcm_distance < 1 || o->glcm_distance > 255))
        return fail(ctx, IMFEAT_ERR_ARG, "glcm_distance must be 1..255");
    for (int k = 0; k < 9; ++k)
        if (!(o->percentiles[k] >= 0.0 && o->percentiles[k] <= 100.0))
            return fail(ctx, IMFEAT_ERR_ARG, "percentiles must lie in [0, 100]");
    return IMFEAT_OK;
}

static void fill_params(Params& P, imfeat_ctx* ctx, const uint16_t* planes, const uint8_t* masks,
                        const int32_t* sizes, const int32_t* src_obj, const int32_t* chan, int64_t n,
                        int c_in, int c_out, int hs, int ws, int64_t plane_stride,
                        const imfeat_opts* o, double* out, int64_t row_stride, uint32_t* status) {
    memset(&P, 0, sizeof(P));
    P.planes = planes; P.masks = masks; P.sizes = sizes; P.src_obj = src_obj; P.chan = chan;
    P.out = out; P.status = status; P.log2tab = ctx->d_log2tab; P.gfix = ctx->d_gfix; P.counts = nullptr;
    P.n_tiles = n * c_out; P.plane_stride = plane_stride; P.row_stride = row_stride;
    P.c_in = c_in; P.c_out = c_out; P.hs = hs; P.ws = ws;
    int col = 0;
    P.col_basic = o->want_basic ? col : -1;   col += o->want_basic ? kNBasic * c_out : 0;
    P.col_glcm = o->want_glcm ? col : -1;     col += o->want_glcm ? kNGlcm * o->n_angles * c_out : 0;
    P.col_shape = o->want_shape ? col : -1;   col += o->want_shape ? kNShape * c_out : 0;
    P.col_moment = o->want_moments ? col : -1;
    P.n_angles = o->want_glcm ? o->n_angles : 0;
    const double PI = 3.14159265358979323846;
    const double ang[4] = {0.0, PI / 4, PI / 2, 3 * PI / 4};
    for (int a = 0; a < kMaxAngles; ++a) {
        P.dr[a] = c_round(sin(ang[a]) * o->glcm_distance);
        P.dc[a] = c_round(cos(ang[a]) * o->glcm_distance);
    }
    for (int k = 0; k < 9; ++k) P.quant[k] = o->percentiles[k] / 100.0;   // numpy: true_divide(q, 100)
}

// Timing ring: slot = {ev[0..4]} recorded around K1,K2,K3,K4 of one extract call; mask bit k set
// when kernel k ran.  Slots are resolved (synchronised + accumulated) lazily, so enabling timing
// does not serialise the stream.
static int timing_resolve(imfeat_ctx* ctx, int slot) {
    const unsigned m = ctx->t_mask[slot];
    if (!m) return IMFEAT_OK;
    static const int kLaunchOrder[4] = {0, 1, 3, 2};       // K1, K2, K4, K3 (see launch_all)
    int last = 0;
    for (int i = 0; i < 4; ++i) if (m & (1u << kLaunchOrder[i])) last = kLaunchOrder[i] + 1;
    CU(cudaEventSynchronize(ctx->t_ev[slot][last]));
    int prev = -1;
    for (int i = 0; i < 4; ++i) {
        const int k = kLaunchOrder[i];
        if (!(m & (1u << k))) continue;
        // the start event of kernel k is the most recent event recorded before it
        int start = (prev < 0) ? 0 : prev + 1;
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, ctx->t_ev[slot][start], ctx->t_ev[slot][k + 1]));
        ctx->t_ms[k] += ms;
        ctx->t_calls[k] += 1;
        prev = k;
    }
    if (m & 16u) {                                         // K4 ran on the side stream, timed by its own events
        float ms = 0.f;
        CU(cudaEventSynchronize(ctx->t_side[slot][1]));
        CU(cudaEventElapsedTime(&ms, ctx->t_side[slot][0], ctx->t_side[slot][1]));
        ctx->t_ms[3] += ms;
        ctx->t_calls[3] += 1;
    }
    ctx->t_mask[slot] = 0;
    return IMFEAT_OK;
}

static int launch_all(imfeat_ctx* ctx, const Params& P_in, const imfeat_opts* o, cudaStream_t st) {
    if (P_in.n_tiles == 0) return IMFEAT_OK;
    Params P = P_in;
    // work counters of this call's dynamically scheduled kernels (ring: calls on different streams
    // may be in flight at the same time)
    P.sched = ctx->d_sched + 8 * (ctx->sched_head++ % kSchedSlots);
    CU(cudaMemsetAsync(P.sched, 0, sizeof(unsigned int) * 8, st));
    P.k1_fp64_only = ctx->env_k1_fp64 != 0;
    int slot = -1;
    if (ctx->timing) {
        slot = ctx->t_head;
        ctx->t_head = (ctx->t_head + 1) % kTimingSlots;
        int rc = timing_resolve(ctx, slot);
        if (rc) return rc;
        for (int k = 0; k < 5; ++k)
            if (!ctx->t_ev[slot][k]) CU(cudaEventCreate(&ctx->t_ev[slot][k]));
        CU(cudaEventRecord(ctx->t_ev[slot][0], st));
    }
#define IMFEAT_MARK(k)                                                  \
    if (slot >= 0) {                                                    \
        CU(cudaEventRecord(ctx->t_ev[slot][(k) + 1], st));              \
        ctx->t_mask[slot] |= 1u << (k);                                 \
    }
    const bool masked = P.masks != nullptr;
    const long long sm = ctx->sm_count;
    int joined = -1;                                       // overlap mode: index of the join event the call ends with
    auto launch_k4w = [&](cudaStream_t s4, long long warps_per_sm) {
        const long long resw = sm * warps_per_sm;
        const int gw = (int)(P.n_tiles < resw ? P.n_tiles : resw);
        const int cpr = P.ws >> 3;
        const bool general = P.sizes != nullptr || (P.ws & 7) != 0 || (cpr & (cpr - 1)) != 0 || cpr > 32;
        if (general) {
            if (masked) k4w_shape_kernel<true, true><<<gw, 32, 0, s4>>>(P);
            else k4w_shape_kernel<false, true><<<gw, 32, 0, s4>>>(P);
        } else {
            if (masked) k4w_shape_kernel<true, false><<<gw, 32, 0, s4>>>(P);
            else k4w_shape_kernel<false, false><<<gw, 32, 0, s4>>>(P);
        }
        ctx->launches += 1;
    };
    const bool k4_warp_tiles = P.hs <= kK4FastDim && P.ws <= kK4FastDim && P.hs * P.ws <= kK4FastPixels && ctx->env_k4_warp != 0;
    const bool k4_overlap = (o->want_shape || o->want_moments) && o->want_glcm && k4_warp_tiles && ctx->env_overlap != 0;
    auto fork_k4 = [&]() -> int {
        // K4w next to the kernels that follow on `st`: a few resident warps per SM on a side stream.  They fill the
        // issue slots the shared-memory-bound GLCM kernels leave free; when they are done their registers go to the
        // CTAs of those kernels that were still waiting for room.
        const unsigned k = ctx->side_head++;
        cudaStream_t s4 = ctx->side[k & 3u];
        if (!s4) { CU(cudaStreamCreateWithFlags(&ctx->side[k & 3u], cudaStreamNonBlocking)); s4 = ctx->side[k & 3u]; }
        cudaEvent_t& fe = ctx->fork_ev[k & 7u];
        cudaEvent_t& je = ctx->join_ev[k & 7u];
        if (!fe) CU(cudaEventCreateWithFlags(&fe, cudaEventDisableTiming));
        if (!je) CU(cudaEventCreateWithFlags(&je, cudaEventDisableTiming));
        CU(cudaEventRecord(fe, st));
        CU(cudaStreamWaitEvent(s4, fe, 0));
        if (slot >= 0) {
            for (int q = 0; q < 2; ++q)
                if (!ctx->t_side[slot][q]) CU(cudaEventCreate(&ctx->t_side[slot][q]));
            CU(cudaEventRecord(ctx->t_side[slot][0], s4));
        }
        launch_k4w(s4, ctx->env_k4_fill);
        if (slot >= 0) {
            CU(cudaEventRecord(ctx->t_side[slot][1], s4));
            ctx->t_mask[slot] |= 16u;
        }
        CU(cudaEventRecord(je, s4));
        joined = (int)(k & 7u);
        return IMFEAT_OK;
    };
    if (k4_overlap && ctx->env_overlap == 2) {             // from the start: next to K12 as well
        int rcf = fork_k4();
        if (rcf) return rcf;
    }
    if (o->want_basic) {
        const int g2 = (int)(P.n_tiles < sm ? P.n_tiles : sm);
        const int ng2 = k2_groups(ctx);
        const bool fuse12 = ctx->env_fuse12 != 0 && ctx->env_k2_compact != 0 && !(ctx->env_k1_tma == 1 && !masked);
        // worklist of the tiles the compact histogram could not take (value range >= 4096): a ring of kWlSlots
        // buffers, because calls on different streams (e.g. the two streams of the host pipeline) may be in flight
        uint32_t* wl = nullptr;
        if (ctx->env_k2_compact != 0) {
            if ((size_t)P.n_tiles + 1 > ctx->worklist_cap) {             // rare: the batch grew
                size_t cap_new = (size_t)P.n_tiles + 1;
                if (cap_new < 2 * ctx->worklist_cap) cap_new = 2 * ctx->worklist_cap;
                ctx->worklist_cap = 0;
                int rcw = grow_buffer(ctx, st, (void**)&ctx->d_worklist, sizeof(uint32_t) * kWlSlots * cap_new, "worklist");
                if (rcw) return rcw;
                ctx->worklist_cap = cap_new;
            }
            wl = ctx->d_worklist + (size_t)(ctx->wl_head++ % kWlSlots) * ctx->worklist_cap;
            CU(cudaMemsetAsync(wl, 0, sizeof(uint32_t), st));
        }
        if (fuse12) {
            // the whole basic block in one pass over the pixels; tiles whose histogram window did not hold are
            // left to the full-range kernel
            const long long res = sm * (ctx->k12_bps[masked] > 0 ? ctx->k12_bps[masked] : 1);
            const int g = (int)(P.n_tiles < res ? P.n_tiles : res);
            if (masked) k12_basic_kernel<true><<<g, 32, 0, st>>>(P, wl + 1, wl);
            else k12_basic_kernel<false><<<g, 32, 0, st>>>(P, wl + 1, wl);
            IMFEAT_MARK(0)
            // the tiles K12 left over: private byte-counter tables first (three CTAs per SM, no hand-over); what
            // wraps a byte counter goes on to K2's 16-bit table through a second list
            uint32_t* wl_k2 = wl;
            if (ctx->env_k2_bytes && ctx->k2b_bps[masked] > 0) {
                wl_k2 = ctx->d_worklist + (size_t)(ctx->wl_head++ % kWlSlots) * ctx->worklist_cap;
                CU(cudaMemsetAsync(wl_k2, 0, sizeof(uint32_t), st));
                const long long resb = sm * ctx->k2b_bps[masked];
                const int gb = (int)(P.n_tiles < resb ? P.n_tiles : resb);
                if (masked) k2b_order_entropy_kernel<true><<<gb, kK2bThreads, sizeof(K2bSmem), st>>>(P, wl + 1, wl, wl_k2 + 1, wl_k2, P.sched + 1);
                else k2b_order_entropy_kernel<false><<<gb, kK2bThreads, sizeof(K2bSmem), st>>>(P, wl + 1, wl, wl_k2 + 1, wl_k2, P.sched + 1);
                ctx->launches += 1;
            }
            if (masked) k2_order_entropy_kernel<true><<<g2, 1024, sizeof(K2Smem), st>>>(P, ng2, wl_k2 + 1, wl_k2);
            else k2_order_entropy_kernel<false><<<g2, 1024, sizeof(K2Smem), st>>>(P, ng2, wl_k2 + 1, wl_k2);
            ctx->launches += 2;
            IMFEAT_MARK(1)
        } else {
            const long long res1 = sm * (ctx->k1_bps[masked] > 0 ? ctx->k1_bps[masked] : 1);   // one resident wave
            const int g1 = (int)((P.n_tiles + 7) / 8 < res1 ? (P.n_tiles + 7) / 8 : res1);
            const bool use_tma = !masked && ctx->env_k1_tma == 1;   // measured: no faster than the direct path
            if (use_tma) {
                // shared-memory ring of whole tiles filled by cp.async.bulk; one persistent CTA per SM
                const int stage_bytes = ((P.hs * P.ws * 2 + 127) & ~127);
                int n_stages = (200 * 1024) / stage_bytes;
                if (n_stages > 32) n_stages = 32;
                const size_t smem = (size_t)n_stages * stage_bytes + 16 * (size_t)n_stages;
                const int gt = (int)(P.n_tiles < sm ? P.n_tiles : sm);
                k1_moments_tma_kernel<<<gt, kK1TmaThreads, smem, st>>>(P, n_stages, stage_bytes);
            } else if (masked) k1_moments_kernel<true><<<g1, 256, 0, st>>>(P);
            else k1_moments_kernel<false><<<g1, 256, 0, st>>>(P);
            IMFEAT_MARK(0)
            if (ctx->env_k2_compact == 0) {
                // full-range kernel for every tile
                if (masked) k2_order_entropy_kernel<true><<<g2, 1024, sizeof(K2Smem), st>>>(P, ng2, nullptr, nullptr);
                else k2_order_entropy_kernel<false><<<g2, 1024, sizeof(K2Smem), st>>>(P, ng2, nullptr, nullptr);
                ctx->launches += 2;
            } else {
                // compact kernel first (value range < 4096, known from K1's min/max); it appends the
                // remaining tiles to the worklist that the full-range ring kernel then works off
                const long long resc = sm * (ctx->k2c_bps[masked] > 0 ? ctx->k2c_bps[masked] : 1);
                const int gc = (int)(P.n_tiles < resc ? P.n_tiles : resc);
                if (masked) k2c_order_entropy_kernel<true><<<gc, kK2cThreads, 0, st>>>(P, wl + 1, wl);
                else k2c_order_entropy_kernel<false><<<gc, kK2cThreads, 0, st>>>(P, wl + 1, wl);
                if (masked) k2_order_entropy_kernel<true><<<g2, 1024, sizeof(K2Smem), st>>>(P, ng2, wl + 1, wl);
                else k2_order_entropy_kernel<false><<<g2, 1024, sizeof(K2Smem), st>>>(P, ng2, wl + 1, wl);
                ctx->launches += 3;
            }
            IMFEAT_MARK(1)
        }
    }
    // K4 goes before K3: with several GPUs the all-gather of the previous batch then overlaps the
    // dynamically scheduled warp-per-tile kernels (K1, K2c, K4w) and is over when K3 starts, whose
    // persistent one-CTA-per-SM grid would otherwise wait for the SMs the collective holds
    if (k4_overlap && joined < 0) {
        // the side stream's work is queued here, behind K12 (IMFEAT_OVERLAP=1), so that it is the first to take
        // its few warp slots when K12 has drained; the GLCM kernels are queued after it on `st`
        int rcf = fork_k4();
        if (rcf) return rcf;
    } else if (!k4_overlap && (o->want_shape || o->want_moments)) {
        if (k4_warp_tiles) {
            // every tile of this batch fits the fast path: one warp per tile, many warps per SM
            launch_k4w(st, ctx->k4w_bps[masked] > 0 ? ctx->k4w_bps[masked] : 1);
            ctx->launches -= 1;                            // counted below
        } else {
            const long long res4 = sm * (ctx->k4_bps[masked] > 0 ? ctx->k4_bps[masked] : 1);
            const int g4 = (int)(P.n_tiles < res4 ? P.n_tiles : res4);
            if (masked) k4_shape_kernel<true><<<g4, kK4Threads, sizeof(K4Smem), st>>>(P);
            else k4_shape_kernel<false><<<g4, kK4Threads, sizeof(K4Smem), st>>>(P);
        }
        IMFEAT_MARK(3)
        ctx->launches += 1;
    }
    if (o->want_glcm) {
        const int maxpx = ((P.hs * P.ws + 7) & ~7);
        if (!masked && ctx->env_k3_ring) {
            // unmasked tiles: the one-kernel ring variant is the faster one (see k3_ring.cuh)
            const int g3 = (int)(P.n_tiles < sm ? P.n_tiles : sm);
            const int ng = ring::k3_groups(maxpx, false);
            const size_t smem3 = ring::k3_smem_bytes(maxpx, false);
            const int pf = ring::k3_prefetch(maxpx, false) ? 1 : 0;
            if (ng == 4) ring::k3_glcm_kernel<false, false, 4><<<g3, ring::kK3Threads, smem3, st>>>(P, maxpx, pf);
            else ring::k3_glcm_kernel<false, false, 2><<<g3, ring::kK3Threads, smem3, st>>>(P, maxpx, pf);
            ctx->launches += 1;
        } else {
            int rc3 = launch_k3<false>(ctx, masked, st, P, k3_q8_capacity(P.hs, P.ws, masked));
            if (rc3) return rc3;
        }
        IMFEAT_MARK(2)
    }
    if (joined >= 0) CU(cudaStreamWaitEvent(st, ctx->join_ev[joined], 0));
#undef IMFEAT_MARK
    CU(cudaGetLastError());
    return IMFEAT_OK;
}

extern "C" {

int imfeat_minmax_fit_device(imfeat_ctx* ctx, const double* d_table, int64_t n_rows, int32_t n_cols,
                             int64_t row_stride, double* d_stats, void* stream) {
    if (!ctx) return fail(nullptr, IMFEAT_ERR_ARG, "ctx is NULL");
    if (n_rows < 0 || n_cols <= 0 || row_stride < n_cols) return fail(ctx, IMFEAT_ERR_ARG, "bad table shape");
    if (!d_stats || (n_rows > 0 && !d_table)) return fail(ctx, IMFEAT_ERR_ARG, "NULL table or stats pointer");
    DeviceGuard guard(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = (int)(n_rows < kPostRowBlocks ? (n_rows > 0 ? n_rows : 1) : kPostRowBlocks);
    double* partial = nullptr;
    CU(cudaMallocAsync((void**)&partial, sizeof(double) * 2 * (size_t)blocks * n_cols, st));
    const dim3 grid((n_cols + 127) / 128, blocks);
    colstats_partial_kernel<<<grid, 128, 0, st>>>(d_table, n_rows, n_cols, row_stride, partial);
    minmax_finish_kernel<<<(n_cols + 127) / 128, 128, 0, st>>>(partial, blocks, n_cols, d_stats);
    ctx->launches += 2;
    CU(cudaGetLastError());
    CU(cudaFreeAsync(partial, st));
    return IMFEAT_OK;
}

int imfeat_minmax_transform_device(imfeat_ctx* ctx, const double* d_in, int64_t n_rows, int32_t n_cols,
                                   int64_t row_stride_in, const double* d_stats, double* d_out,
                                   int64_t row_stride_out, void* stream) {
    if (!ctx) return fail(nullptr, IMFEAT_ERR_ARG, "ctx is NULL");
    if (n_rows < 0 || n_cols <= 0 || row_stride_in < n_cols || row_stride_out < n_cols)
        return fail(ctx, IMFEAT_ERR_ARG, "bad table shape");
    if (n_rows == 0) return IMFEAT_OK;
    if (!d_in || !d_out || !d_stats) return fail(ctx, IMFEAT_ERR_ARG, "NULL pointer");
    DeviceGuard guard(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)n_rows * n_cols;
    const long long want = (total + 255) / 256, cap = 8ll * ctx->sm_count;
    minmax_apply_kernel<<<(int)(want < cap ? want : cap), 256, 0, st>>>(d_in, n_rows, n_cols, row_stride_in, d_stats,
                                                                       d_out, row_stride_out);
    ctx->launches += 1;
    CU(cudaGetLastError());
    return IMFEAT_OK;
}

int imfeat_enable_timing(imfeat_ctx* ctx, int32_t enable) {
    if (!ctx) return fail(nullptr, IMFEAT_ERR_ARG, "ctx is NULL");
    ctx->timing = enable ? 1 : 0;
    return IMFEAT_OK;
}

int imfeat_kernel_times(imfeat_ctx* ctx, double* ms_out, int64_t* calls_out, int32_t reset) {
    if (!ctx) return fail(nullptr, IMFEAT_ERR_ARG, "ctx is NULL");
    DeviceGuard guard(ctx->device);
    for (int sl = 0; sl < kTimingSlots; ++sl) {
        int rc = timing_resolve(ctx, sl);
        if (rc) return rc;
    }
    for (int k = 0; k < 4; ++k) {
        if (ms_out) ms_out[k] = ctx->t_ms[k];
        if (calls_out) calls_out[k] = ctx->t_calls[k];
        if (reset) { ctx->t_ms[k] = 0.0; ctx->t_calls[k] = 0; }
    }
    return IMFEAT_OK;
}

int imfeat_extract_device(imfeat_ctx* ctx, const uint16_t* d_planes, const uint8_t* d_masks,
                          const int32_t* d_sizes, const int32_t* d_src_obj, const int32_t* d_chan,
                          int64_t n_objects, int32_t c_in, int32_t c_out, int32_t hs, int32_t ws,
                          int64_t plane_stride, const imfeat_opts* opts, double* d_out,
                          int64_t row_stride, uint32_t* d_status, void* stream) {
    int rc = check_common(ctx, d_planes, n_objects, c_in, c_out, hs, ws, plane_stride, opts);
    if (rc) return rc;
    if (!d_chan && c_out != c_in)
        return fail(ctx, IMFEAT_ERR_ARG, "c_out (%d) != c_in (%d) needs a channel list", c_out, c_in);
    if (n_objects > 0 && !d_out) return fail(ctx, IMFEAT_ERR_ARG, "d_out is NULL");
    if (row_stride < imfeat_row_width(c_out, opts))
        return fail(ctx, IMFEAT_ERR_ARG, "row_stride %lld < row width %lld", (long long)row_stride,
                    (long long)imfeat_row_width(c_out, opts));
    if (d_masks && ((uintptr_t)d_masks & 7)) return fail(ctx, IMFEAT_ERR_ARG, "masks must be 8-byte aligned");
    if (n_objects == 0) return IMFEAT_OK;
    DeviceGuard guard(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (d_status) { CU(cudaMemsetAsync(d_status, 0, sizeof(uint32_t) * n_objects, st)); }
    Params P;
    fill_params(P, ctx, d_planes, d_masks, d_sizes, d_src_obj, d_chan, n_objects, c_in, c_out, hs,
                ws, plane_stride, opts, d_out, row_stride, d_status);
    return launch_all(ctx, P, opts, st);
}

int imfeat_glcm_counts_device(imfeat_ctx* ctx, const uint16_t* d_planes, const uint8_t* d_masks,
                              const int32_t* d_sizes, int64_t n_objects, int32_t c, int32_t hs,
                              int32_t ws, int64_t plane_stride, const imfeat_opts* opts,
                              uint32_t* d_counts, void* stream) {
    int rc = check_common(ctx, d_planes, n_objects, c, c, hs, ws, plane_stride, opts);
    if (rc) return rc;
    if (!opts->want_glcm) return fail(ctx, IMFEAT_ERR_ARG, "opts->want_glcm must be set");
    if (n_objects > 0 && !d_counts) return fail(ctx, IMFEAT_ERR_ARG, "d_counts is NULL");
    if (n_objects == 0) return IMFEAT_OK;
    DeviceGuard guard(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    // the property columns are computed too; park them in a scratch table
    imfeat_opts o = *opts;
    o.want_basic = 0; o.want_shape = 0; o.want_moments = 0;
    const int64_t width = imfeat_row_width(c, &o);
    double* scratch = nullptr;
    CU(cudaMallocAsync((void**)&scratch, sizeof(double) * width * n_objects, st));
    Params P;
    fill_params(P, ctx, d_planes, d_masks, d_sizes, nullptr, nullptr, n_objects, c, c, hs, ws,
                plane_stride, &o, scratch, width, nullptr);
    P.counts = d_counts;
    const bool masked = d_masks != nullptr;
    const int maxpx = k3_q8_capacity(P.hs, P.ws, masked);
    // the tile counter of K3's dynamic scheduler (slot 2 of a work-counter set)
    P.sched = ctx->d_sched + 8 * (ctx->sched_head++ % kSchedSlots);
    CU(cudaMemsetAsync(P.sched, 0, sizeof(unsigned int) * 8, st));
    rc = launch_k3<true>(ctx, masked, st, P, maxpx);
    if (rc) return rc;
    CU(cudaGetLastError());
    CU(cudaFreeAsync(scratch, st));
    return IMFEAT_OK;
}

int imfeat_pack_hwc_device(imfeat_ctx* ctx, const uint16_t* d_hwc, const uint8_t* d_mask_hwc,
                           const int32_t* d_sizes, int64_t n_objects, int32_t c, int32_t hs,
                           int32_t ws, int64_t plane_stride, uint16_t* d_planes, uint8_t* d_masks,
                           void* stream) {
    if (!ctx) return fail(nullptr, IMFEAT_ERR_ARG, "ctx is NULL");
    if (n_objects < 0 || c < 1 || hs < 1 || ws < 1 || plane_stride < (int64_t)hs * ws)
        return fail(ctx, IMFEAT_ERR_ARG, "bad shape");
    if (n_objects == 0) return IMFEAT_OK;
    if (!d_hwc || !d_planes) return fail(ctx, IMFEAT_ERR_ARG, "NULL image pointer");
    if ((d_mask_hwc == nullptr) != (d_masks == nullptr))
        return fail(ctx, IMFEAT_ERR_ARG, "mask input and output must both be given or both be NULL");
    DeviceGuard guard(ctx->device);
    const long long total = (long long)n_objects * hs * ws;
    const long long want = (total + 255) / 256;
    const int grid = (int)(want < (long long)ctx->sm_count * 16 ? want : (long long)ctx->sm_count * 16);
    pack_hwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_hwc, d_mask_hwc, d_sizes, n_objects, c, hs,
                                                           ws, plane_stride, d_planes, d_masks, 0);
    ctx->launches += 1;
    CU(cudaGetLastError());
    return IMFEAT_OK;
}

int imfeat_synth_device(imfeat_ctx* ctx, uint64_t seed, int64_t first_object, int64_t n_objects,
                        int32_t c, int32_t hs, int32_t ws, int64_t plane_stride, int32_t variable,
                        int32_t hmin, int32_t wmin, int32_t mask_shrink_256, uint16_t* d_planes,
                        uint8_t* d_masks, int32_t* d_sizes, void* stream) {
    if (!ctx) return fail(nullptr, IMFEAT_ERR_ARG, "ctx is NULL");
    if (n_objects < 0 || c < 1 || hs < 1 || ws < 1 || plane_stride < (int64_t)hs * ws)
        return fail(ctx, IMFEAT_ERR_ARG, "bad shape");
    if (variable && (hmin < 1 || hmin > hs || wmin < 1 || wmin > ws || !d_sizes))
        return fail(ctx, IMFEAT_ERR_ARG, "variable sizes need 1 <= hmin <= hs, 1 <= wmin <= ws and d_sizes");
    if (n_objects == 0) return IMFEAT_OK;
    if (!d_planes) return fail(ctx, IMFEAT_ERR_ARG, "d_planes is NULL");
    DeviceGuard guard(ctx->device);
    const long long planes = (long long)n_objects * c;
    const int grid = (int)(planes < (long long)ctx->sm_count * 32 ? planes : (long long)ctx->sm_count * 32);
    synth_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(seed, first_object, n_objects, c, hs, ws,
                                                        plane_stride, variable, hmin, wmin,
                                                        mask_shrink_256, d_planes, d_masks, d_sizes);
    ctx->launches += 1;
    CU(cudaGetLastError());
    return IMFEAT_OK;
}

// ---------------------------------------------------------------------------------------------
// Host-buffer entry point: slab pipeline  (pinned staging -> H2D -> kernels -> D2H), two slabs in
// flight on two streams so copies overlap the kernels of the neighbouring slab.
// ---------------------------------------------------------------------------------------------
static bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// layout: 0 = plane-compact planar uint16[N][c][plane_stride]; 1 = interleaved README.md:8 layout
// uint16[N][hs][ws][c] (valid region = top-left h_i x w_i), packed to planar on the device.
static int extract_host_impl(imfeat_ctx* ctx, int layout, const uint16_t* h_img, const uint8_t* h_masks,
                             const int32_t* h_sizes, int64_t n_objects, int32_t c, int32_t hs,
                             int32_t ws, int64_t plane_stride, const imfeat_opts* opts, double* h_out,
                             int64_t row_stride, uint32_t* h_status) {
    int rc = check_common(ctx, h_img, n_objects, c, c, hs, ws, plane_stride, opts, false);
    if (rc) return rc;
    if (n_objects > 0 && !h_out) return fail(ctx, IMFEAT_ERR_ARG, "h_out is NULL");
    if (h_sizes)                                           // a size beyond the stride would read other objects' pixels
        for (int64_t i = 0; i < n_objects; ++i) {
            const int32_t h = h_sizes[2 * i], w = h_sizes[2 * i + 1];
            if (h < 1 || w < 1 || h > hs || w > ws || (int64_t)h * w > plane_stride)
                return fail(ctx, IMFEAT_ERR_ARG, "object %lld: size %dx%d outside 1..%dx%d", (long long)i, h, w, hs, ws);
        }
    const int64_t width = imfeat_row_width(c, opts);
    if (row_stride < width) return fail(ctx, IMFEAT_ERR_ARG, "row_stride < row width");
    if (n_objects == 0) return IMFEAT_OK;
    DeviceGuard guard(ctx->device);
    for (int b = 0; b < 2; ++b) {
        if (!ctx->streams[b]) CU(cudaStreamCreateWithFlags(&ctx->streams[b], cudaStreamNonBlocking));
        if (!ctx->done[b]) CU(cudaEventCreateWithFlags(&ctx->done[b], cudaEventDisableTiming));
    }
    const size_t src_px = layout ? (size_t)hs * ws * c : (size_t)c * plane_stride;  // host elems/object
    const size_t pl_px = (size_t)c * plane_stride;                                  // planar elems/object
    // host mask bytes per object: one per element, or bit-packed (every object / plane padded to 8 bytes)
    const bool mbits = h_masks && opts->host_mask_bits != 0;
    const size_t src_mk = !h_masks ? 0 : !mbits ? src_px
                        : layout ? (size_t)IMFEAT_MASK_BITS_BYTES(src_px) : (size_t)c * (size_t)IMFEAT_MASK_BITS_BYTES(plane_stride);
    const bool planar_copy_px = layout != 0;                 // interleaved pixels are packed to planes on the device
    const bool planar_copy_mk = h_masks && (layout != 0 || mbits);
    const size_t obj_in = src_px * 2 + src_mk + 8 + (planar_copy_px ? pl_px * 2 : 0) + (planar_copy_mk ? pl_px : 0);
    const size_t obj_out = (size_t)width * 8 + 4;
    int64_t slab = (int64_t)(((size_t)64 << 20) / (src_px * 2));      // ~64 MiB of pixels per slab
    if (slab < 1) slab = 1;
    if (slab > n_objects) slab = n_objects;
    const size_t need_in = (size_t)slab * obj_in + 256, need_out = (size_t)slab * obj_out + 64;
    if (need_in > ctx->in_bytes || need_out > ctx->out_bytes) {
        // nobody else may still use the old staging buffers: the two host-path streams are drained first
        for (int b = 0; b < 2; ++b) CU(cudaStreamSynchronize(ctx->streams[b]));
        free_staging(ctx);
        for (int b = 0; b < 2; ++b) {
            CU(cudaMallocHost(&ctx->pin_in[b], need_in));
            CU(cudaMallocHost(&ctx->pin_out[b], need_out));
            CU(cudaMalloc(&ctx->dev_in[b], need_in));
            CU(cudaMalloc(&ctx->dev_out[b], need_out));
        }
        ctx->in_bytes = need_in;
        ctx->out_bytes = need_out;
    }
    const bool direct_in = is_pinned(h_img) && (!h_masks || is_pinned(h_masks));
    // results go straight into the caller's buffer when it is pinned and dense (no staging copy on the host)
    const bool direct_out = row_stride == width && is_pinned(h_out) && (!h_status || is_pinned(h_status));
    const int64_t n_slabs = (n_objects + slab - 1) / slab;
    // byte offsets inside a slab buffer: [pixels | masks | sizes | planar pixels | planar masks]
    auto up16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const size_t off_mask = up16((size_t)slab * src_px * 2);
    const size_t off_size = up16(off_mask + (size_t)slab * src_mk);
    const size_t off_plpx = up16(off_size + (size_t)slab * 8);
    const size_t off_plmk = up16(off_plpx + (planar_copy_px ? (size_t)slab * pl_px * 2 : 0));
    const size_t off_stat = (size_t)slab * width * 8;
    auto drain = [&](int64_t s) {   // copy slab s's results from pinned staging to the caller
        if (direct_out) return;
        const int b = (int)(s & 1);
        const int64_t first = s * slab, cnt = (first + slab <= n_objects) ? slab : n_objects - first;
        const double* src = (const double*)ctx->pin_out[b];
        if (row_stride == width) memcpy(h_out + first * row_stride, src, (size_t)cnt * width * 8);
        else for (int64_t i = 0; i < cnt; ++i) memcpy(h_out + (first + i) * row_stride, src + i * width, (size_t)width * 8);
        if (h_status) memcpy(h_status + first, (const char*)ctx->pin_out[b] + off_stat, (size_t)cnt * 4);
    };
    for (int64_t s = 0; s < n_slabs; ++s) {
        const int b = (int)(s & 1);
        cudaStream_t st = ctx->streams[b];
        const int64_t first = s * slab, cnt = (first + slab <= n_objects) ? slab : n_objects - first;
        if (s >= 2) { CU(cudaEventSynchronize(ctx->done[b])); drain(s - 2); }
        char* din = (char*)ctx->dev_in[b];
        char* pin = (char*)ctx->pin_in[b];
        const uint16_t* src_img = h_img + (size_t)first * src_px;
        const uint8_t* src_mask = h_masks ? h_masks + (size_t)first * src_mk : nullptr;
        if (direct_in) {
            CU(cudaMemcpyAsync(din, src_img, (size_t)cnt * src_px * 2, cudaMemcpyHostToDevice, st));
            if (h_masks) CU(cudaMemcpyAsync(din + off_mask, src_mask, (size_t)cnt * src_mk, cudaMemcpyHostToDevice, st));
        } else {
            memcpy(pin, src_img, (size_t)cnt * src_px * 2);
            CU(cudaMemcpyAsync(din, pin, (size_t)cnt * src_px * 2, cudaMemcpyHostToDevice, st));
            if (h_masks) {
                memcpy(pin + off_mask, src_mask, (size_t)cnt * src_mk);
                CU(cudaMemcpyAsync(din + off_mask, pin + off_mask, (size_t)cnt * src_mk, cudaMemcpyHostToDevice, st));
            }
        }
        if (h_sizes) {
            memcpy(pin + off_size, h_sizes + 2 * first, (size_t)cnt * 8);
            CU(cudaMemcpyAsync(din + off_size, pin + off_size, (size_t)cnt * 8, cudaMemcpyHostToDevice, st));
        }
        const int32_t* d_sizes = h_sizes ? (const int32_t*)(din + off_size) : nullptr;
        const uint16_t* d_planes = (const uint16_t*)din;
        const uint8_t* d_masks = h_masks ? (const uint8_t*)(din + off_mask) : nullptr;
        if (layout) {
            const long long total = (long long)cnt * hs * ws, want = (total + 255) / 256;
            const int grid = (int)(want < (long long)ctx->sm_count * 16 ? want : (long long)ctx->sm_count * 16);
            pack_hwc_kernel<<<grid, 256, 0, st>>>((const uint16_t*)din, d_masks, d_sizes, cnt, c, hs, ws, plane_stride,
                                                  (uint16_t*)(din + off_plpx), h_masks ? (uint8_t*)(din + off_plmk) : nullptr,
                                                  mbits ? 1 : 0);
            ctx->launches += 1;
            d_planes = (const uint16_t*)(din + off_plpx);
            d_masks = h_masks ? (const uint8_t*)(din + off_plmk) : nullptr;
        } else if (mbits) {
            const long long total = (long long)cnt * c * (plane_stride >> 3), want = (total + 255) / 256;
            const int grid = (int)(want < (long long)ctx->sm_count * 16 ? want : (long long)ctx->sm_count * 16);
            unpack_mask_bits_kernel<<<grid, 256, 0, st>>>(d_masks, (long long)cnt * c, plane_stride, (uint8_t*)(din + off_plmk));
            ctx->launches += 1;
            d_masks = (const uint8_t*)(din + off_plmk);
        }
        double* dout = (double*)ctx->dev_out[b];
        uint32_t* dstat = (uint32_t*)((char*)ctx->dev_out[b] + off_stat);
        CU(cudaMemsetAsync(dstat, 0, (size_t)cnt * 4, st));
        Params P;
        fill_params(P, ctx, d_planes, d_masks, d_sizes, nullptr, nullptr, cnt, c, c, hs, ws, plane_stride,
                    opts, dout, width, dstat);
        rc = launch_all(ctx, P, opts, st);
        if (rc) return rc;
        if (direct_out) {
            CU(cudaMemcpyAsync(h_out + first * row_stride, dout, (size_t)cnt * width * 8, cudaMemcpyDeviceToHost, st));
            if (h_status) CU(cudaMemcpyAsync(h_status + first, dstat, (size_t)cnt * 4, cudaMemcpyDeviceToHost, st));
        } else {
            CU(cudaMemcpyAsync(ctx->pin_out[b], dout, (size_t)cnt * width * 8, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync((char*)ctx->pin_out[b] + off_stat, dstat, (size_t)cnt * 4, cudaMemcpyDeviceToHost, st));
        }
        CU(cudaEventRecord(ctx->done[b], st));
    }
    for (int64_t s = (n_slabs >= 2 ? n_slabs - 2 : 0); s < n_slabs; ++s) {
        CU(cudaEventSynchronize(ctx->done[s & 1]));
        drain(s);
    }
    return IMFEAT_OK;
}

int imfeat_extract_host(imfeat_ctx* ctx, const uint16_t* h_planes, const uint8_t* h_masks,
                        const int32_t* h_sizes, int64_t n_objects, int32_t c, int32_t hs, int32_t ws,
                        int64_t plane_stride, const imfeat_opts* opts, double* h_out,
                        int64_t row_stride, uint32_t* h_status) {
    return extract_host_impl(ctx, 0, h_planes, h_masks, h_sizes, n_objects, c, hs, ws, plane_stride, opts,
                             h_out, row_stride, h_status);
}

int imfeat_extract_host_hwc(imfeat_ctx* ctx, const uint16_t* h_hwc, const uint8_t* h_mask_hwc,
                            const int32_t* h_sizes, int64_t n_objects, int32_t c, int32_t hs, int32_t ws,
                            const imfeat_opts* opts, double* h_out, int64_t row_stride,
                            uint32_t* h_status) {
    const int64_t plane_stride = ((int64_t)hs * ws + 7) & ~(int64_t)7;
    return extract_host_impl(ctx, 1, h_hwc, h_mask_hwc, h_sizes, n_objects, c, hs, ws, plane_stride, opts,
                             h_out, row_stride, h_status);
}

}  // extern "C"
