#pragma once
// kfusion::utils::Quaternion<T> -- interface-compatible with the reference's kfusion/src/utils/quaternion.hpp (same member
// names w_/x_/y_/z_, same methods) so user code and the reference's tests compile against it.  Written for this repo; the
// arithmetic keeps the reference's evaluation order (float products, DOUBLE scale factor in normalize()) because the GPU
// kernels (csrc/warp_common.cuh) and the oracle are pinned bit for bit to that order (tests/golden/dq_ref.json).
#include <cassert>
#include <cmath>
#include <iostream>
#include <kfusion/types.hpp>

namespace kfusion
{
    namespace utils
    {
        template <typename T> class Quaternion
        {
        public:
            T w_, x_, y_, z_;

            Quaternion() : w_(1), x_(0), y_(0), z_(0) {}
            Quaternion(T w, T x, T y, T z) : w_(w), x_(x), y_(y), z_(z) {}

            /** rotation taking the z axis onto `normal` (frame completed with two tangents) */
            Quaternion(const Vec3f& normal)
            {
                Vec3f t0 = normal.cross(Vec3f(1, 0, 0));
                if (t0.dot(t0) < 0.001f) t0 = normal.cross(Vec3f(0, 1, 0));
                t0 = cv::normalize(t0);
                Vec3f t1 = cv::normalize(normal.cross(t0));
                const float m[3][3] = {{t0[0], t0[1], t0[2]}, {t1[0], t1[1], t1[2]}, {normal[0], normal[1], normal[2]}};
                w_ = std::sqrt(1.0 + m[0][0] + m[1][1] + m[2][2]) / 2.0;
                x_ = (m[2][1] - m[1][2]) / (w_ * 4);
                y_ = (m[0][2] - m[2][0]) / (w_ * 4);
                z_ = (m[1][0] - m[2][1]) / (w_ * 4);
                if (norm() > 0) normalize();
            }

            /** unit quaternion of a rotation by theta (radians) about (x, y, z) */
            void encodeRotation(T theta, T x, T y, T z)
            {
                const T s = std::sin(theta / 2);
                w_ = std::cos(theta / 2);
                x_ = x * s; y_ = y * s; z_ = z * s;
                normalize();
            }

            void getRodrigues(T& x, T& y, T& z)
            {
                if (w_ == 1) { x = y = z = 0; return; }
                const T half_theta = std::acos(w_);
                const T k = std::sin(half_theta) * std::tan(half_theta);
                x = x_ / k; y = y_ / k; z = z_ / k;
            }

            /** q (0,v) q*  -- no normalisation */
            void rotate(T& x, T& y, T& z)
            {
                Quaternion<T> q = *this;
                Quaternion<T> r = q * Quaternion<T>(0, x, y, z) * q.conjugate();
                x = r.x_; y = r.y_; z = r.z_;
            }

            /** v += 2 u x (u x v + w v) with (w, u) the normalised copy of *this */
            void rotate(Vec3f& v) const
            {
                Quaternion<T> r = *this;
                r.normalize();
                const Vec3f u(r.x_, r.y_, r.z_);
                v += (u * 2.f).cross(u.cross(v) + v * r.w_);
            }

            Quaternion operator+(const Quaternion& o) { return Quaternion(w_ + o.w_, x_ + o.x_, y_ + o.y_, z_ + o.z_); }
            void operator+=(const Quaternion& o) { *this = *this + o; }
            Quaternion operator-(const Quaternion& o) { return Quaternion(w_ - o.w_, x_ - o.x_, y_ - o.y_, z_ - o.z_); }
            Quaternion operator-() { return Quaternion(-w_, -x_, -y_, -z_); }
            bool operator==(const Quaternion& o) const { return w_ == o.w_ && x_ == o.x_ && y_ == o.y_ && z_ == o.z_; }

            template <typename U> friend Quaternion operator*(const U s, const Quaternion& q) { return Quaternion<T>(s * q.w_, s * q.x_, s * q.y_, s * q.z_); }
            template <typename U> friend Quaternion operator/(const Quaternion& q, const U s) { return (1 / s) * q; }

            /** Hamilton product */
            Quaternion operator*(const Quaternion& o)
            {
                return Quaternion((w_ * o.w_) - (x_ * o.x_) - (y_ * o.y_) - (z_ * o.z_),
                                  (w_ * o.x_) + (x_ * o.w_) + (y_ * o.z_) - (z_ * o.y_),
                                  (w_ * o.y_) - (x_ * o.z_) + (y_ * o.w_) + (z_ * o.x_),
                                  (w_ * o.z_) + (x_ * o.y_) - (y_ * o.x_) + (z_ * o.w_));
            }

            T dotProduct(Quaternion o) { return 0.5 * ((conjugate() * o) + (*this) * o.conjugate()).w_; }
            Quaternion conjugate() const { return Quaternion<T>(w_, -x_, -y_, -z_); }
            T norm() { return std::sqrt((w_ * w_) + (x_ * x_) + (y_ * y_) + (z_ * z_)); }

            void normalize()
            {
                assert(!((w_ == 0) && (x_ == 0) && (y_ == 0) && (z_ == 0)));
                const T n = norm();
                assert(n > 0);
                *this = (1.0 / n) * (*this);          // double scale factor, narrowed per component
            }

            template <typename U> friend std::ostream& operator<<(std::ostream& os, const Quaternion<U>& q)
            { return os << "(" << q.w_ << ", " << q.x_ << ", " << q.y_ << ", " << q.z_ << ")"; }
        };
    }
}
