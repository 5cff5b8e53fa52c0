// kfref_host.cpp -- TEST INFRASTRUCTURE (oracle/_ref/libkfref.so).  The reference keeps five trivial host-side definitions
// of its device-layer structs in translation units that need OpenCV (kfusion/src/precomp.cpp:24-55,
// kfusion/src/projective_icp.cpp:11-23).  They are member-initialisers only and are restated here so the reference's CUDA
// sources link; everything arithmetic comes from the reference's own files.
#include <cmath>
#include "internal.hpp"

kfusion::device::TsdfVolume::TsdfVolume(elem_type *d, int3 dm, float3 vs, float td, int mw) : data(d), dims(dm), voxel_size(vs), trunc_dist(td), max_weight(mw) {}
kfusion::device::Projector::Projector(float fx, float fy, float cx, float cy) : f(make_float2(fx, fy)), c(make_float2(cx, cy)) {}
kfusion::device::Reprojector::Reprojector(float fx, float fy, float cx, float cy) : finv(make_float2(1.f / fx, 1.f / fy)), c(make_float2(cx, cy)) {}
kfusion::device::ComputeIcpHelper::ComputeIcpHelper(float dist_thres, float angle_thres)
{
    min_cosine = cos(angle_thres);
    dist2_thres = dist_thres * dist_thres;
}
void kfusion::device::ComputeIcpHelper::setLevelIntr(int level_index, float fx, float fy, float cx, float cy)
{
    int div = 1 << level_index;
    f = make_float2(fx / div, fy / div);
    c = make_float2(cx / div, cy / div);
    finv = make_float2(1.f / f.x, 1.f / f.y);
}
