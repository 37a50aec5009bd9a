// ref_utils_shim.cpp -- C entry point onto the REFERENCE's own CPU top-K
// comparator, topKsort<float, unsigned int> (U/Utils.cpp:213-243, explicitly
// instantiated at U/Utils.cpp:245-246), which the reference's TestSort.cpp uses
// as the truth for kCalculateTopK.  Linked with the unmodified Utils.cpp into
// oracle/_ref/libdsstne_refutils.so.  TEST INFRASTRUCTURE ONLY.
#include "Utils.h"

// topKsort<> is declared by Utils.h:117.

extern "C" void ref_topKsort_f32_u32(float* keys, unsigned int* vals, int size, float* topKkeys,
                                     unsigned int* topKvals, int topK)
{
    topKsort<float, unsigned int>(keys, vals, size, topKkeys, topKvals, topK, true);
}
