// knn_ref -- runs the REFERENCE's k-NN: the vendored nanoflann (kfusion/include/nanoflann/nanoflann.hpp) with the reference's
// point-cloud adaptor (kfusion/src/utils/knn_point_cloud.hpp), both compiled from /root/reference as they lie, configured
// exactly as WarpField does (warp_field.cpp:20-24 leaf size 10, KNNResultSet(8); :247-251 SearchParams(10)).
// stdin:  P Q, then P+Q points (x y z);  stdout: per query 8 x "index dist_hex".
#include <cstdio>
#include <nanoflann/nanoflann.hpp>
#include <knn_point_cloud.hpp>

typedef nanoflann::KDTreeSingleIndexAdaptor<nanoflann::L2_Simple_Adaptor<float, kfusion::utils::PointCloud>, kfusion::utils::PointCloud, 3> kd_tree_t;

int main()
{
    int P, Q;
    if (scanf("%d %d", &P, &Q) != 2) return 1;
    kfusion::utils::PointCloud cloud;
    cloud.pts.resize(P);
    for (int i = 0; i < P; ++i) { double x, y, z; if (scanf("%lf %lf %lf", &x, &y, &z) != 3) return 1; cloud.pts[i] = cv::Vec3f((float)x, (float)y, (float)z); }
    kd_tree_t index(3, cloud, nanoflann::KDTreeSingleIndexAdaptorParams(10));
    index.buildIndex();
    std::vector<size_t> ret_index(8);
    std::vector<float> out_dist_sqr(8);
    nanoflann::KNNResultSet<float> resultSet(8);
    for (int q = 0; q < Q; ++q) {
        double x, y, z;
        if (scanf("%lf %lf %lf", &x, &y, &z) != 3) return 1;
        cv::Vec3f p((float)x, (float)y, (float)z);
        resultSet.init(&ret_index[0], &out_dist_sqr[0]);
        index.findNeighbors(resultSet, p.val, nanoflann::SearchParams(10));
        for (int i = 0; i < 8; ++i) printf("%zu %a ", ret_index[i], (double)out_dist_sqr[i]);
        printf("\n");
    }
    return 0;
}
