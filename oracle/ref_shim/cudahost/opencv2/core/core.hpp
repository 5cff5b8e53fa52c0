// oracle/ref_shim/cudahost: the reference's device layer includes this header but uses nothing from it
#pragma once
