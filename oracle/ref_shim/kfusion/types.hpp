// ref_shim/kfusion/types.hpp -- stands in for the reference's kfusion/types.hpp (which pulls OpenCV core/affine/viz, absent in
// this image) so that the reference's OWN headers -- kfusion/src/utils/{quaternion,dual_quaternion,knn_point_cloud}.hpp and the
// vendored nanoflann -- compile unmodified from where they lie under /root/reference.  It only provides the handful of
// cv:: types those headers name, with OpenCV 2.4's arithmetic (float ops, left-to-right).  Written for this repo; not a
// copy of OpenCV or of the reference.
#pragma once
#include <cassert>   // the reference's headers rely on OpenCV having pulled these in
#include <cmath>
#include <cstddef>
#include <iostream>
#include <vector>

namespace cv
{
    struct Vec3f
    {
        float val[3];
        Vec3f() { val[0] = val[1] = val[2] = 0.f; }
        Vec3f(float a, float b, float c) { val[0] = a; val[1] = b; val[2] = c; }
        float& operator[](int i) { return val[i]; }
        const float& operator[](int i) const { return val[i]; }
        Vec3f cross(const Vec3f& v) const
        { return Vec3f(val[1] * v.val[2] - val[2] * v.val[1], val[2] * v.val[0] - val[0] * v.val[2], val[0] * v.val[1] - val[1] * v.val[0]); }
        float dot(const Vec3f& v) const { return val[0] * v.val[0] + val[1] * v.val[1] + val[2] * v.val[2]; }
        Vec3f& operator+=(const Vec3f& o) { val[0] += o.val[0]; val[1] += o.val[1]; val[2] += o.val[2]; return *this; }
    };
    inline Vec3f operator+(const Vec3f& a, const Vec3f& b) { return Vec3f(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
    inline Vec3f operator-(const Vec3f& a, const Vec3f& b) { return Vec3f(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
    inline Vec3f operator*(const Vec3f& a, float s) { return Vec3f(a[0] * s, a[1] * s, a[2] * s); }
    inline Vec3f operator*(float s, const Vec3f& a) { return Vec3f(a[0] * s, a[1] * s, a[2] * s); }
    inline bool operator!=(const Vec3f& a, const Vec3f& b) { return a[0] != b[0] || a[1] != b[1] || a[2] != b[2]; }
    inline bool operator==(const Vec3f& a, const Vec3f& b) { return !(a != b); }
    inline std::ostream& operator<<(std::ostream& os, const Vec3f& v) { return os << "[" << v[0] << ", " << v[1] << ", " << v[2] << "]"; }
    inline Vec3f normalize(const Vec3f& v) { float n = std::sqrt(v.dot(v)); return Vec3f(v[0] / n, v[1] / n, v[2] / n); }

    // only named by Quaternion(const Vec3f& normal), which the pinned paths never instantiate
    struct Mat3f
    {
        std::vector<Vec3f> rows;
        void push_back(const Vec3f& r) { rows.push_back(r); }
        template <typename T> T at(int r, int c) const { return (T)rows[r][c]; }
    };
}

namespace kfusion
{
    typedef cv::Vec3f Vec3f;
}
