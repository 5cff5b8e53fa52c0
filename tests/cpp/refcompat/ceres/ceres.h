// stand-in: tests/ceres_warp_test.cpp of the reference includes ceres/ceres.h but only drives kfusion::WarpField (energy_data is served by the
// device LM/PCG solver here; Ceres is absent from this image)
