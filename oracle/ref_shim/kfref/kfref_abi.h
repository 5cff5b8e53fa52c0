// kfref_abi.h -- TEST INFRASTRUCTURE.  C ABI of oracle/_ref/libkfref.so: the REFERENCE's own kernels
// (kfusion/src/cuda/*.cu compiled for the host through oracle/ref_shim/cudahost) behind exactly the signatures of the
// oracle's restatement (oracle/orc_*.c), so a test can run both on the same buffers and compare bit for bit.
#pragma once
#include <stddef.h>
#include <stdint.h>
typedef struct { uint32_t *data; int dims[3]; float voxel_size[3]; float trunc_dist; int max_weight; } kfref_volume;   // = orc_volume
typedef struct { float R[9]; float t[3]; } kfref_aff3f;                                                               // = orc_aff3f
typedef struct { float fx, fy, cx, cy; } kfref_intr;                                                                  // = orc_intr
