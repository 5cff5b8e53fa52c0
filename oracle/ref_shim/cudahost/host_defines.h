// host stand-in (oracle/ref_shim/cudahost): everything lives in cuda_runtime.h
#pragma once
#include "cuda_runtime.h"
