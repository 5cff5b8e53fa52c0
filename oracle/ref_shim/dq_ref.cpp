// dq_ref -- runs the REFERENCE's quaternion / dual-quaternion classes (kfusion/src/utils/quaternion.hpp, dual_quaternion.hpp,
// compiled from /root/reference as they lie) through the blend-and-transform sequence of WarpField::DQB / warp
// (warp_field.cpp:180-217): used to pin the oracle's restated arithmetic (double scalars, normalisation order) bit for bit.
// stdin: K (number of cases); per case: 8 x { rot w x y z, trans x y z, weight }, then point x y z.
// stdout per case: blended rotation (4 hex floats), transformed point (3 hex floats).
#include <cstdio>
#include <dual_quaternion.hpp>
using namespace kfusion::utils;

int main()
{
    int K;
    if (scanf("%d", &K) != 1) return 1;
    for (int c = 0; c < K; ++c) {
        Quaternion<float> translation_sum(0, 0, 0, 0), rotation_sum(0, 0, 0, 0);
        for (int i = 0; i < 8; ++i) {
            double rw, rx, ry, rz, tx, ty, tz, w;
            if (scanf("%lf %lf %lf %lf %lf %lf %lf %lf", &rw, &rx, &ry, &rz, &tx, &ty, &tz, &w) != 8) return 1;
            // node transform built as the reference builds it: DualQuaternion(Quaternion(0,t), rotation)
            DualQuaternion<float> node(Quaternion<float>(0, (float)tx, (float)ty, (float)tz), Quaternion<float>((float)rw, (float)rx, (float)ry, (float)rz));
            float weight = (float)w;
            translation_sum += weight * node.getTranslation();          // warp_field.cpp:211
            rotation_sum += weight * node.getRotation();                // :212
        }
        rotation_sum.normalize();                                       // :214
        DualQuaternion<float> res(translation_sum, rotation_sum);       // :215
        double px, py, pz;
        if (scanf("%lf %lf %lf", &px, &py, &pz) != 3) return 1;
        cv::Vec3f point((float)px, (float)py, (float)pz);
        res.transform(point);                                           // :187
        Quaternion<float> r = res.getRotation();
        printf("%a %a %a %a %a %a %a\n", (double)r.w_, (double)r.x_, (double)r.y_, (double)r.z_, (double)point[0], (double)point[1], (double)point[2]);
    }
    return 0;
}
