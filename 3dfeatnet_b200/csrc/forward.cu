// forward.cu -- C ABI of the fused detector / descriptor forward; dispatches on `precision`.
#include "common.cuh"
#include "weights_layout.h"

namespace f3d {
int detector_forward_fp32(int b, int n, int m, int S, float radius, const float *xyz, const float *new_xyz, const int *idx,
                          const float *packed, float *pooled_ws, float *attention, float *orientation, cudaStream_t st);
int descriptor_forward_fp32(int b, int n, int m, int S, float radius, int feature_dim, const float *xyz,
                            const float *new_xyz, const int *idx, const float *orientation, const float *packed,
                            float *pooled_ws, float *features, cudaStream_t st);
int detector_rows_tc(long long num_clusters, int n, int m, float radius, const float *xyz, const float *new_xyz,
                     const int *idx, const float *packed, uint8_t *wimg, float *pooled, bool build_image, int max_ctas, cudaStream_t st);
int descriptor_rows_tc(long long num_clusters, int n, int m, float radius, int feature_dim, const float *xyz,
                       const float *new_xyz, const int *idx, const float *orientation, const float *packed, uint8_t *wimg,
                       float *pooledA, float *pmaxh, bool build_image, int max_ctas, cudaStream_t st);
size_t descriptor_tc_weight_bytes();
int descriptor_post_fp32(long long nc, const float *pooled2, const float *packed, int feature_dim, float *features,
                         cudaStream_t st);
size_t post_tc_weight_bytes();
int detector_post_tc(long long nc, const float *pooled, const float *packed, uint8_t *wimg, float *attention, float *orientation,
                     bool build_image, int max_ctas, cudaStream_t st);
int descriptor_post_tc(long long nc, int feature_dim, const float *pooledA, const float *pmaxh, const float *packed, uint8_t *wimg, float *features,
                       bool build_image, int max_ctas, cudaStream_t st);
int detector_post_fp32(long long num_clusters, const float *pooled, const float *packed, float *attention,
                       float *orientation, cudaStream_t st);
}  // namespace f3d
extern "C" size_t f3d_detector_tc_weight_bytes(void);

using namespace f3d;

static size_t pooled_bytes(int b, int m) { return (static_cast<size_t>(b) * m * 256 * sizeof(float) + 255) & ~static_cast<size_t>(255); }

F3D_API size_t f3d_packed_weights_floats(int feature_dim) { return static_cast<size_t>(make_weight_layout(feature_dim).total); }
F3D_API int f3d_packed_weights_num_blocks(void) { return kNumWeightSlots; }
F3D_API int f3d_packed_weights_offsets(int feature_dim, int *offsets, int *sizes) {
    if (!offsets || !sizes) return fail(F3D_ERR_INVALID_ARGUMENT, "packed_weights_offsets: null output");
    const WeightLayout L = make_weight_layout(feature_dim);
    for (int i = 0; i < kNumWeightSlots; ++i) {
        offsets[i] = L.off[i];
        sizes[i] = L.size[i];
    }
    return 0;
}

// Workspace = pooled vectors | four weight-image slots (detector rows, detector tail, descriptor rows, descriptor tail).  The
// slots do not alias, so the images survive from call to call and F3D_PRECISION_IMAGES_CACHED can skip rebuilding them.
static size_t slot_bytes(int slot) {
    const size_t b = slot == 0 ? f3d_detector_tc_weight_bytes() : (slot == 2 ? descriptor_tc_weight_bytes() : post_tc_weight_bytes());
    return (b + 255) & ~static_cast<size_t>(255);
}
static uint8_t *image_slot(void *workspace, int b, int m, int slot) {
    uint8_t *p = static_cast<uint8_t *>(workspace) + pooled_bytes(b, m);
    for (int s = 0; s < slot; ++s) p += slot_bytes(s);
    return p;
}
F3D_API size_t f3d_forward_workspace_bytes(int b, int m, int feature_dim) {
    (void)feature_dim;
    return pooled_bytes(b, m) + slot_bytes(0) + slot_bytes(1) + slot_bytes(2) + slot_bytes(3) + 256;
}

F3D_API int f3d_detector_forward(int b, int n, int m, int nsample, float radius, const float *xyz, const float *new_xyz,
                                 const int *idx, const float *packed, float *attention, float *orientation,
                                 int precision, void *workspace, size_t workspace_bytes, void *stream) {
    if (b < 0 || n <= 0 || m < 0 || nsample <= 0 || !(radius > 0.0f) || !xyz || !new_xyz || !idx || !packed || !attention ||
        !orientation)
        return fail(F3D_ERR_INVALID_ARGUMENT, "detector_forward: bad arguments");
    if (!workspace || workspace_bytes < f3d_forward_workspace_bytes(b, m, 32))
        return fail(F3D_ERR_WORKSPACE_TOO_SMALL, "detector_forward: workspace too small");
    const bool build_images = (precision & F3D_PRECISION_IMAGES_CACHED) == 0;
    const int max_ctas = (precision >> 16) & 0xff;  // F3D_PRECISION_SM_LIMIT(n): SMs granted to the persistent kernels (0 = all)
    precision &= 0xff;
    if (precision == 0)
        return detector_forward_fp32(b, n, m, nsample, radius, xyz, new_xyz, idx, packed, static_cast<float *>(workspace),
                                     attention, orientation, as_stream(stream));
    if (precision == 2) {  // tcgen05, bf16x3 split
        if (nsample != 64) return fail(F3D_ERR_UNSUPPORTED, "detector_forward: the tensor-core path needs nsample == 64");
        float *pooled = static_cast<float *>(workspace);
        int rc = detector_rows_tc(static_cast<long long>(b) * m, n, m, radius, xyz, new_xyz, idx, packed, image_slot(workspace, b, m, 0), pooled,
                                  build_images, max_ctas, as_stream(stream));
        if (rc) return rc;
        return detector_post_tc(static_cast<long long>(b) * m, pooled, packed, image_slot(workspace, b, m, 1), attention, orientation,
                                build_images, max_ctas, as_stream(stream));
    }
    return fail(F3D_ERR_UNSUPPORTED, "detector_forward: precision must be 0 (fp32) or 2 (bf16x3 tensor cores)");
}

F3D_API int f3d_descriptor_forward(int b, int n, int m, int nsample, float radius, int feature_dim, const float *xyz,
                                   const float *new_xyz, const int *idx, const float *orientation, const float *packed,
                                   float *features, int precision, void *workspace, size_t workspace_bytes,
                                   void *stream) {
    if (b < 0 || n <= 0 || m < 0 || nsample <= 0 || !(radius > 0.0f) || !xyz || !new_xyz || !idx || !packed || !features)
        return fail(F3D_ERR_INVALID_ARGUMENT, "descriptor_forward: bad arguments");
    if (!workspace || workspace_bytes < f3d_forward_workspace_bytes(b, m, feature_dim))
        return fail(F3D_ERR_WORKSPACE_TOO_SMALL, "descriptor_forward: workspace too small");
    const bool build_images = (precision & F3D_PRECISION_IMAGES_CACHED) == 0;
    const int max_ctas = (precision >> 16) & 0xff;  // F3D_PRECISION_SM_LIMIT(n)
    precision &= 0xff;
    if (precision == 0)
        return descriptor_forward_fp32(b, n, m, nsample, radius, feature_dim, xyz, new_xyz, idx, orientation, packed,
                                       static_cast<float *>(workspace), features, as_stream(stream));
    if (precision == 2 && nsample == 64 && feature_dim <= 64) {  // tcgen05, bf16x3 split
        // the row kernel leaves, per cluster, the per-point half of conv_mid_0 max-pooled (128 floats) and the two sample-half maxima of
        // the 64 conv1 channels (2 x 64 floats); the tail adds the pooled half of conv_mid_0 (the workspace holds 256 floats per cluster)
        const long long nc = static_cast<long long>(b) * m;
        float *pooledA = static_cast<float *>(workspace);
        float *pmaxh = pooledA + nc * 128;
        int rc = descriptor_rows_tc(nc, n, m, radius, feature_dim, xyz, new_xyz, idx, orientation, packed, image_slot(workspace, b, m, 2),
                                    pooledA, pmaxh, build_images, max_ctas, as_stream(stream));
        if (rc) return rc;
        return descriptor_post_tc(nc, feature_dim, pooledA, pmaxh, packed, image_slot(workspace, b, m, 3), features, build_images, max_ctas,
                                  as_stream(stream));
    }
    if (precision == 2)  // shapes the tensor-core kernel does not cover (nsample != 64, feature_dim 128): exact fp32 kernel
        return descriptor_forward_fp32(b, n, m, nsample, radius, feature_dim, xyz, new_xyz, idx, orientation, packed,
                                       static_cast<float *>(workspace), features, as_stream(stream));
    return fail(F3D_ERR_UNSUPPORTED, "descriptor_forward: precision must be 0 (fp32) or 2 (bf16x3 tensor cores)");
}

// ---- output rows of inference.py ----------------------------------------------------------------------------------------
namespace f3d {
__global__ void pack_rows_kernel(long long rows, int feature_dim, const float *__restrict__ xyz, const float *__restrict__ attention,
                                 const float *__restrict__ orientation, const float *__restrict__ features, float *__restrict__ out) {
    const int cols = 5 + feature_dim;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const long long r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    float v;
    if (c < 3) v = __ldg(xyz + r * 3 + c);
    else if (c == 3) v = __ldg(attention + r);
    else if (c == 4) v = __ldg(orientation + r);
    else v = __ldg(features + r * feature_dim + (c - 5));
    out[i] = v;
}
}  // namespace f3d

// One row [x y z | attention | orientation | descriptor] per keypoint: the layout the host side of the end-to-end path copies
// back in one transfer (inference.py:174-177 writes [xyz | descriptor] rows from the same fields).
F3D_API int f3d_pack_rows(long long rows, int feature_dim, const float *xyz, const float *attention, const float *orientation,
                          const float *features, float *out, void *stream) {
    if (rows < 0 || feature_dim <= 0 || !xyz || !attention || !orientation || !features || !out)
        return fail(F3D_ERR_INVALID_ARGUMENT, "pack_rows: bad arguments");
    const long long total = rows * (5 + feature_dim);
    if (total == 0) return 0;
    pack_rows_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, as_stream(stream)>>>(rows, feature_dim, xyz, attention, orientation,
                                                                                                  features, out);
    return check_launch("pack_rows_kernel");
}
