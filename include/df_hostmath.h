/*
 * df_hostmath.h -- host-side float affine helpers shared by the pipeline (dynamicfusion_b200/csrc/pipeline.cu), the
 * C++ mirror classes and the CPU oracle's pipeline restatement, so that both sides derive bit-identical vol2cam /
 * cam2vol / Rinv / pose chains from the same inputs.
 *
 * They restate what the reference does on the host with OpenCV 2.4.13 (absent here, "parity unpinned" for these
 * few 3x3 operations): cv::Affine3f::inv (tsdf_volume.cpp:112,135; kinfu.cpp:356), Affine3f * Affine3f
 * (kinfu.cpp:280), Matx33f::inv (tsdf_volume.cpp:136,165,215), Affine3f * Vec3f (kinfu.cpp:361).
 * Plain C, float arithmetic, fixed evaluation order (compile with -ffp-contract=off / -fmad=false).
 */
#ifndef DF_HOSTMATH_H
#define DF_HOSTMATH_H

#ifdef __cplusplus
extern "C" {
#endif

/* R (row-major 3x3) inverse by cofactors / determinant, float (cv::Matx33f::inv closed form) */
static inline void dfh_mat3_inv(const float *R, float *out)
{
    const float a = R[0], b = R[1], c = R[2], d = R[3], e = R[4], f = R[5], g = R[6], h = R[7], i = R[8];
    const float c00 = e * i - f * h, c01 = f * g - d * i, c02 = d * h - e * g;
    const float det = a * c00 + b * c01 + c * c02;
    const float id = 1.f / det;
    out[0] = c00 * id; out[1] = (c * h - b * i) * id; out[2] = (b * f - c * e) * id;
    out[3] = c01 * id; out[4] = (a * i - c * g) * id; out[5] = (c * d - a * f) * id;
    out[6] = c02 * id; out[7] = (b * g - a * h) * id; out[8] = (a * e - b * d) * id;
}

/* (R, t)^-1 = (R^-1, -R^-1 t); aff = 12 floats: R row-major then t */
static inline void dfh_aff_inv(const float *aff, float *out)
{
    dfh_mat3_inv(aff, out);
    for (int r = 0; r < 3; ++r)
        out[9 + r] = -(out[r * 3 + 0] * aff[9] + out[r * 3 + 1] * aff[10] + out[r * 3 + 2] * aff[11]);
}

/* out = A * B */
static inline void dfh_aff_mul(const float *A, const float *B, float *out)
{
    float tmp[12];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c)
            tmp[r * 3 + c] = A[r * 3 + 0] * B[0 * 3 + c] + A[r * 3 + 1] * B[1 * 3 + c] + A[r * 3 + 2] * B[2 * 3 + c];
        tmp[9 + r] = A[r * 3 + 0] * B[9] + A[r * 3 + 1] * B[10] + A[r * 3 + 2] * B[11] + A[9 + r];
    }
    for (int k = 0; k < 12; ++k) out[k] = tmp[k];
}

static inline void dfh_aff_identity(float *out)
{
    for (int k = 0; k < 12; ++k) out[k] = 0.f;
    out[0] = out[4] = out[8] = 1.f;
}

#ifdef __cplusplus
}
#endif
#endif
