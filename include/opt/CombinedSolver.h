#pragma once
// The reference's kfusion/include/opt/CombinedSolver.h wraps the Opt (Terra) solver.  Here the class keeps its constructor
// and entry points (initializeProblemInstance / solveAll) but runs the device LM/PCG of libdfusion.so
// (df_solve_data_term); CombinedSolverParameters keeps the reference's field names
// (deps/Opt/examples/shared/CombinedSolverParameters.h:3-15).
#include <vector>
#include <kfusion/warp_field.hpp>

struct CombinedSolverParameters
{
    bool useCUDA = false;
    bool useOpt = true;
    bool useOptLM = false;
    bool useCeres = false;
    bool earlyOut = false;
    unsigned int numIter = 1;
    unsigned int nonLinearIter = 3;
    unsigned int linearIter = 200;
    unsigned int patchIter = 32;
    bool profileSolve = true;
    std::string optDoublePrecision = "false";
};

class CombinedSolver
{
public:
    CombinedSolver(kfusion::WarpField *warpField, CombinedSolverParameters params);
    ~CombinedSolver();
    void initializeProblemInstance(const std::vector<cv::Vec3f> &canonical_vertices, const std::vector<cv::Vec3f> &canonical_normals,
                                   const std::vector<cv::Vec3f> &live_vertices, const std::vector<cv::Vec3f> &live_normals);
    void solveAll();
    double lastCost() const { return last_cost_; }
private:
    kfusion::WarpField *m_warp;
    CombinedSolverParameters m_combinedSolverParameters;
    struct Impl;
    Impl *impl_;
    double last_cost_ = 0.0;
};
