// imfeat_torch.cpp -- thin PyTorch C++ extension over the C ABI of include/imfeat.h (libimfeat.so).
//
// Registers the torch ops SURVEY.md 8(b) sketches, so that tensors, the current CUDA stream and CUDA-graph capture go
// through torch natively:
//   imfeat::extract(planes, masks?, sizes?, src_obj?, chan?, hs, ws, basic, glcm, n_angles, distance, shape, moments,
//                   percentiles, out?, status?, ctx) -> Tensor f64[N, row_width]          (replaces NB:358-364 for one batch)
//   imfeat::glcm_counts(planes, masks?, sizes?, hs, ws, n_angles, distance, ctx) -> Tensor i32[N, C, A, 256, 256]   (NB:298 bins)
//   imfeat::row_width(c_out, basic, glcm, n_angles, shape, moments) -> int
// All arithmetic stays in libimfeat.so's kernels; this file only checks tensors and forwards pointers.  `ctx` is the
// address of an imfeat_ctx to use (FeatureExtractor passes its own, so that its timing and work buffers apply), or 0
// for a context this library keeps per device.
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <mutex>
#include <vector>

#include "../../include/imfeat.h"

namespace {

std::mutex g_mu;
std::vector<imfeat_ctx*> g_ctx;            // per device, created on first use

imfeat_ctx* context_for(int64_t handle, int device) {
    if (handle != 0) return reinterpret_cast<imfeat_ctx*>(static_cast<intptr_t>(handle));
    std::lock_guard<std::mutex> lock(g_mu);
    if ((int)g_ctx.size() <= device) g_ctx.resize(device + 1, nullptr);
    if (!g_ctx[device]) {
        const int rc = imfeat_create(device, &g_ctx[device]);
        TORCH_CHECK(rc == IMFEAT_OK, "imfeat_create failed: ", imfeat_last_error(nullptr));
    }
    return g_ctx[device];
}

imfeat_opts make_opts(bool basic, bool glcm, int64_t n_angles, int64_t distance, bool shape, bool moments,
                      c10::ArrayRef<double> percentiles) {
    imfeat_opts o;
    imfeat_default_opts(&o);
    o.want_basic = basic; o.want_glcm = glcm; o.n_angles = (int32_t)n_angles; o.glcm_distance = (int32_t)distance;
    o.want_shape = shape; o.want_moments = moments;
    if (!percentiles.empty()) {
        TORCH_CHECK(percentiles.size() == 9, "exactly nine percentile arguments (NB:242-250)");
        for (int k = 0; k < 9; ++k) o.percentiles[k] = percentiles[k];
    }
    return o;
}

const void* opt_ptr(const c10::optional<at::Tensor>& t, at::ScalarType st, const at::Tensor& like, const char* name) {
    if (!t.has_value() || !t->defined()) return nullptr;
    TORCH_CHECK(t->is_cuda() && t->device() == like.device(), name, " must live on the device of planes");
    TORCH_CHECK(t->scalar_type() == st && t->is_contiguous(), name, ": wrong dtype or not contiguous");
    return t->data_ptr();
}

void check_planes(const at::Tensor& planes) {
    TORCH_CHECK(planes.is_cuda() && planes.is_contiguous() && planes.dim() == 3, "planes: contiguous cuda [N, C, plane_stride]");
    TORCH_CHECK(planes.element_size() == 2, "planes must be 16-bit");
}

at::Tensor extract(const at::Tensor& planes, const c10::optional<at::Tensor>& masks, const c10::optional<at::Tensor>& sizes,
                   const c10::optional<at::Tensor>& src_obj, const c10::optional<at::Tensor>& chan, int64_t hs, int64_t ws,
                   bool basic, bool glcm, int64_t n_angles, int64_t distance, bool shape, bool moments,
                   c10::ArrayRef<double> percentiles, const c10::optional<at::Tensor>& out, const c10::optional<at::Tensor>& status,
                   int64_t ctx_handle) {
    check_planes(planes);
    const c10::cuda::CUDAGuard guard(planes.device());
    const int64_t N = planes.size(0), C = planes.size(1), stride = planes.size(2);
    const imfeat_opts o = make_opts(basic, glcm, n_angles, distance, shape, moments, percentiles);
    const void* d_masks = nullptr;
    if (masks.has_value() && masks->defined()) {
        TORCH_CHECK(masks->is_cuda() && masks->is_contiguous() && masks->element_size() == 1 && masks->numel() == planes.numel(),
                    "masks: contiguous 8-bit cuda tensor of the shape of planes");
        d_masks = masks->data_ptr();
    }
    const bool has_chan = chan.has_value() && chan->defined();
    const int64_t c_out = has_chan ? chan->numel() : C;
    const int64_t width = imfeat_row_width((int32_t)c_out, &o);
    at::Tensor table = (out.has_value() && out->defined()) ? *out : at::empty({N, width}, planes.options().dtype(at::kDouble));
    TORCH_CHECK(table.is_cuda() && table.scalar_type() == at::kDouble && table.dim() == 2 && table.size(0) == N &&
                table.size(1) >= width && table.stride(1) == 1, "out: cuda float64 [N, >= row_width], unit column stride");
    imfeat_ctx* ctx = context_for(ctx_handle, planes.device().index());
    const int rc = imfeat_extract_device(
        ctx, static_cast<const uint16_t*>(planes.data_ptr()), static_cast<const uint8_t*>(d_masks),
        static_cast<const int32_t*>(opt_ptr(sizes, at::kInt, planes, "sizes")),
        static_cast<const int32_t*>(opt_ptr(src_obj, at::kInt, planes, "src_obj")),
        static_cast<const int32_t*>(opt_ptr(chan, at::kInt, planes, "chan")), N, (int32_t)C, (int32_t)c_out, (int32_t)hs,
        (int32_t)ws, stride, &o, table.data_ptr<double>(), table.stride(0),
        static_cast<uint32_t*>(const_cast<void*>(opt_ptr(status, at::kInt, planes, "status"))),
        c10::cuda::getCurrentCUDAStream(planes.device().index()).stream());
    TORCH_CHECK(rc == IMFEAT_OK, "imfeat_extract_device: ", imfeat_last_error(ctx));
    return table;
}

at::Tensor glcm_counts(const at::Tensor& planes, const c10::optional<at::Tensor>& masks, const c10::optional<at::Tensor>& sizes,
                       int64_t hs, int64_t ws, int64_t n_angles, int64_t distance, int64_t ctx_handle) {
    check_planes(planes);
    const c10::cuda::CUDAGuard guard(planes.device());
    const int64_t N = planes.size(0), C = planes.size(1), stride = planes.size(2);
    const imfeat_opts o = make_opts(false, true, n_angles, distance, false, false, {});
    const void* d_masks = nullptr;
    if (masks.has_value() && masks->defined()) {
        TORCH_CHECK(masks->is_cuda() && masks->is_contiguous() && masks->element_size() == 1 && masks->numel() == planes.numel(),
                    "masks: contiguous 8-bit cuda tensor of the shape of planes");
        d_masks = masks->data_ptr();
    }
    at::Tensor counts = at::empty({N, C, n_angles, 256, 256}, planes.options().dtype(at::kInt));
    imfeat_ctx* ctx = context_for(ctx_handle, planes.device().index());
    const int rc = imfeat_glcm_counts_device(
        ctx, static_cast<const uint16_t*>(planes.data_ptr()), static_cast<const uint8_t*>(d_masks),
        static_cast<const int32_t*>(opt_ptr(sizes, at::kInt, planes, "sizes")), N, (int32_t)C, (int32_t)hs, (int32_t)ws, stride,
        &o, reinterpret_cast<uint32_t*>(counts.data_ptr<int32_t>()), c10::cuda::getCurrentCUDAStream(planes.device().index()).stream());
    TORCH_CHECK(rc == IMFEAT_OK, "imfeat_glcm_counts_device: ", imfeat_last_error(ctx));
    return counts;
}

int64_t row_width(int64_t c_out, bool basic, bool glcm, int64_t n_angles, bool shape, bool moments) {
    const imfeat_opts o = make_opts(basic, glcm, n_angles, 5, shape, moments, {});
    return imfeat_row_width((int32_t)c_out, &o);
}

}  // namespace

TORCH_LIBRARY(imfeat, m) {
    m.def("extract(Tensor planes, Tensor? masks, Tensor? sizes, Tensor? src_obj, Tensor? chan, int hs, int ws, bool basic, "
          "bool glcm, int n_angles, int distance, bool shape, bool moments, float[] percentiles, Tensor(a!)? out, "
          "Tensor(b!)? status, int ctx=0) -> Tensor");
    m.def("glcm_counts(Tensor planes, Tensor? masks, Tensor? sizes, int hs, int ws, int n_angles, int distance, int ctx=0) -> Tensor");
    m.def("row_width(int c_out, bool basic, bool glcm, int n_angles, bool shape, bool moments) -> int", &row_width);
}

TORCH_LIBRARY_IMPL(imfeat, CUDA, m) {
    m.impl("extract", &extract);
    m.impl("glcm_counts", &glcm_counts);
}
