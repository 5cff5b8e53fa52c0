#ifndef KFUSION_WARP_FIELD_HPP
#define KFUSION_WARP_FIELD_HPP
// kfusion::WarpField -- same interface as the reference's kfusion/warp_field.hpp:35-88.  The node table also lives on the
// device ([M][12] floats, include/dfusion.h) and is re-synchronised from getNodes() before every device call, so callers
// may edit nodes through the returned vector exactly as with the reference.  The nanoflann kd-tree (kd_tree_t typedef of
// the reference header) is replaced by the device node grid; KNN() still fills getRetIndex() / getDistSquared().
#include <dual_quaternion.hpp>
#include <kfusion/types.hpp>
#include <kfusion/cuda/tsdf_volume.hpp>

#define KNN_NEIGHBOURS 8
namespace kfusion
{
    struct deformation_node
    {
        Vec3f vertex;
        kfusion::utils::DualQuaternion<float> transform;
        float weight = 0;
    };

    class WarpField
    {
    public:
        WarpField();
        ~WarpField();

        void init(const cv::Mat& first_frame);
        void init(const std::vector<Vec3f>& first_frame);
        void energy(const cuda::Cloud &frame, const cuda::Normals &normals, const Affine3f &pose, const cuda::TsdfVolume &tsdfVolume,
                    const std::vector<std::pair<kfusion::utils::DualQuaternion<float>, kfusion::utils::DualQuaternion<float>>> &edges);
        void energy_data(const std::vector<Vec3f> &canonical_vertices, const std::vector<Vec3f> &canonical_normals,
                         const std::vector<Vec3f> &live_vertices, const std::vector<Vec3f> &live_normals);
        void energy_reg(const std::vector<std::pair<kfusion::utils::DualQuaternion<float>, kfusion::utils::DualQuaternion<float>>> &edges);

        void warp(std::vector<Vec3f>& points, std::vector<Vec3f>& normals) const;
        utils::DualQuaternion<float> DQB(const Vec3f& vertex) const;
        void getWeightsAndUpdateKNN(const Vec3f& vertex, float weights[KNN_NEIGHBOURS]) const;
        float weighting(float squared_dist, float weight) const;
        void KNN(Vec3f point) const;
        void clear();

        const std::vector<deformation_node>* getNodes() const;
        std::vector<deformation_node>* getNodes();
        const cv::Mat getNodesAsMat() const;
        void setWarpToLive(const Affine3f &pose);
        std::vector<float>* getDistSquared() const;
        std::vector<size_t>* getRetIndex() const;
        void buildKDTree();

        // not in the reference ("Extending the warp field - stubbed out functionality", Report.md): append a node, made as init()
        // makes them, for every step-th point of the canonical cloud (1 x N CV_32FC4, e.g. TsdfVolume::get_cloud_host()) whose nearest
        // node is farther than radius -> df_extend_field.  Returns the new node count.
        int extend(const cv::Mat& canonical_cloud, float radius, int step = 50, int max_nodes = 4096);

        // device side (not in the reference): used by KinFu / WarpFieldOptimiser
        void uploadNodes() const;                    // host vector -> device table (+ node grid when vertices changed)
        void downloadTranslations();                 // device table -> host vector (after the solve)
        float *deviceNodes() const;
        void *deviceGrid() const;
        int deviceNodeCount() const;
        const Affine3f& getWarpToLive() const { return warp_to_live_; }
        void adoptDeviceNodes(float *nodes_dev, void *grid_dev, int M);   // KinFu: share the pipeline's table
    private:
        std::vector<deformation_node>* nodes_;
        struct Impl;
        Impl* impl_;
        Affine3f warp_to_live_;
    };
}
#endif //KFUSION_WARP_FIELD_HPP
