#pragma once
// kfusion::utils::DualQuaternion<T> -- interface-compatible with the reference's kfusion/src/utils/dual_quaternion.hpp:
// rotation_ is an ordinary rotation quaternion, translation_ the dual part 0.5 * (0, t) * rotation_.  Written for this repo;
// evaluation order follows the reference (see quaternion.hpp in this directory).
#include <cmath>
#include <iostream>
#include <utility>
#include <quaternion.hpp>

namespace kfusion
{
    namespace utils
    {
        static float epsilon() { return 1e-6; }

        template <typename T> class DualQuaternion
        {
        public:
            DualQuaternion() : rotAngle_(0) { rotation_ = Quaternion<float>(); translation_ = Quaternion<float>(); }
            ~DualQuaternion() {}

            /** position + Euler angles (roll, pitch, yaw) */
            DualQuaternion(T x, T y, T z, T roll, T pitch, T yaw) : rotAngle_(0)
            {
                const T cr = std::cos(roll / 2), sr = std::sin(roll / 2);
                const T cp = std::cos(pitch / 2), sp = std::sin(pitch / 2);
                const T cy = std::cos(yaw / 2), sy = std::sin(yaw / 2);
                rotation_.w_ = cr * cp * cy + sr * sp * sy;
                rotation_.x_ = sr * cp * cy - cr * sp * sy;
                rotation_.y_ = cr * sp * cy + sr * cp * sy;
                rotation_.z_ = cr * cp * sy - sr * sp * cy;
                translation_ = 0.5 * Quaternion<T>(0, x, y, z) * rotation_;
            }

            /** translation given as the pure quaternion (0, x, y, z) */
            DualQuaternion(Quaternion<T> translation, Quaternion<T> rotation) : rotAngle_(0)
            {
                rotation_ = rotation;
                translation_ = 0.5 * translation * rotation;
            }

            void encodeRotation(T angle, T x, T y, T z) { rotation_.encodeRotation(angle, x, y, z); }
            void encodeRotation(T x, T y, T z) { rotation_.encodeRotation(std::sqrt(x * x + y * y + z * z), x, y, z); }
            void encodeTranslation(T x, T y, T z) { translation_ = 0.5 * Quaternion<T>(0, x, y, z) * rotation_; }

            /** re-normalise the rotation, keeping the translation */
            void normalize()
            {
                T x, y, z;
                getTranslation(x, y, z);
                rotation_.normalize();
                encodeTranslation(x, y, z);
            }

            void getTranslation(T& x, T& y, T& z) const
            {
                const Quaternion<T> t = getTranslation();
                x = t.x_; y = t.y_; z = t.z_;
            }
            void getTranslation(Vec3f& v) const { getTranslation(v[0], v[1], v[2]); }
            Quaternion<T> getTranslation() const
            {
                Quaternion<T> rot = rotation_;
                rot.normalize();
                return 2 * translation_ * rot.conjugate();
            }

            void getEuler(T& roll, T& pitch, T& yaw) { roll = getRoll(); pitch = getPitch(); yaw = getYaw(); }
            Quaternion<T> getRotation() const { return rotation_; }

            DualQuaternion operator+(const DualQuaternion& o)
            { DualQuaternion r; r.rotation_ = rotation_ + o.rotation_; r.translation_ = translation_ + o.translation_; return r; }
            DualQuaternion operator-(const DualQuaternion& o)
            { DualQuaternion r; r.rotation_ = rotation_ - o.rotation_; r.translation_ = translation_ - o.translation_; return r; }
            DualQuaternion operator*(const DualQuaternion& o)
            { DualQuaternion<T> r; r.rotation_ = rotation_ * o.rotation_; r.translation_ = translation_ + o.translation_; return r; }
            DualQuaternion operator/(const std::pair<T, T> divisor)
            { DualQuaternion<T> r; r.rotation_ = 1 / divisor.first * rotation_; r.translation_ = 1 / divisor.second * translation_; return r; }
            template <typename U> friend DualQuaternion operator*(const U s, const DualQuaternion& q)
            { DualQuaternion<T> r; r.rotation_ = s * q.rotation_; r.translation_ = s * q.translation_; return r; }

            DualQuaternion conjugate()
            { DualQuaternion<T> r; r.rotation_ = rotation_.conjugate(); r.translation_ = translation_.conjugate(); return r; }
            inline DualQuaternion identity() { return DualQuaternion(Quaternion<T>(0, 0, 0, 0), Quaternion<T>(0, 1, 0, 0)); }

            /** rotate, then translate */
            void transform(Vec3f& point)
            {
                Vec3f t;
                getTranslation(t);
                rotation_.rotate(point);
                point += t;
            }

            void from_twist(const float& r0, const float& r1, const float& r2, const float& x, const float& y, const float& z)
            {
                const float n = std::sqrt(r0 * r0 + r1 * r1 + r2 * r2);
                Quaternion<T> rot;
                if (n > epsilon()) {
                    float c = std::cos(n);
                    const float sign = (c > 0.f) - (c < 0.f);
                    c *= sign;
                    const float s_over_n = sign * std::sin(n) / n;
                    rot = Quaternion<T>(c, r0 * s_over_n, r1 * s_over_n, r2 * s_over_n);
                }
                *this = DualQuaternion<T>(Quaternion<T>(0, x, y, z), rot);
            }

            std::pair<T, T> magnitude()
            {
                DualQuaternion r = (*this) * (*this).conjugate();
                return std::make_pair(r.rotation_.w_, r.translation_.w_);
            }

        private:
            Quaternion<T> rotation_;
            Quaternion<T> translation_;
            T position_[3] = {};
            T rotAxis_[3] = {};
            T rotAngle_;

            T getRoll()
            { return std::atan2(2 * ((rotation_.w_ * rotation_.x_) + (rotation_.y_ * rotation_.z_)), (1 - 2 * ((rotation_.x_ * rotation_.x_) + (rotation_.y_ * rotation_.y_)))); }
            T getPitch() { return std::asin(2 * (rotation_.w_ * rotation_.y_ - rotation_.z_ * rotation_.x_)); }
            T getYaw()
            { return std::atan2(2 * ((rotation_.w_ * rotation_.z_) + (rotation_.x_ * rotation_.y_)), (1 - 2 * ((rotation_.y_ * rotation_.y_) + (rotation_.z_ * rotation_.z_)))); }
        };

        template <typename T> std::ostream& operator<<(std::ostream& os, const DualQuaternion<T>& q)
        { return os << "[" << q.getRotation() << ", " << q.getTranslation() << ", " << "]" << std::endl; }
    }
}
