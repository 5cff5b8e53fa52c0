// ply_check.cpp -- CPU-only check of include/kfusion/io/ply.hpp: writes a cloud the way get_cloud_host() lays it out
#include <kfusion/io/ply.hpp>
#include <cstdio>
#include <limits>
int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    const int n = 7;
    cv::Mat cloud(1, n, CV_32FC4), normals(1, n, CV_32FC4);
    for (int i = 0; i < n; ++i) {
        float *p = cloud.ptr<float>() + 4 * i, *q = normals.ptr<float>() + 4 * i;
        p[0] = 0.5f * i; p[1] = -1.f * i; p[2] = 2.f + i; p[3] = 0.f;
        q[0] = 0.f; q[1] = 0.f; q[2] = 1.f; q[3] = 0.f;
    }
    cloud.ptr<float>()[4 * 3] = std::numeric_limits<float>::quiet_NaN();        // point 3 is dropped
    normals.ptr<float>()[4 * 5 + 1] = std::numeric_limits<float>::quiet_NaN();  // normal 5 is written as 0 0 0
    const long a = kfusion::writePly(std::string(argv[1]) + "/with_normals.ply", cloud, normals);
    const long b = kfusion::writePly(std::string(argv[1]) + "/points_only.ply", cloud);
    const long c = kfusion::writePly(std::string(argv[1]) + "/no/such/dir/x.ply", cloud);
    std::printf("%ld %ld %ld\n", a, b, c);
    return 0;
}
