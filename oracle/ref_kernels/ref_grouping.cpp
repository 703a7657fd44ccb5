// Test-only wrapper around the REFERENCE's own C++ hand grouping (src/cpp_grouping/grouping.cpp), compiled unchanged from
// where it lies under /root/reference (found through -I, see oracle/Makefile target `ref`).  TEST INFRASTRUCTURE - never linked
// into the product.  The reference exposes it through Cython (cpp_grouping.pyx); this exports the same call as plain C.
#include <grouping.cpp>   // reference: src/cpp_grouping/grouping.cpp

extern "C" __attribute__((visibility("default")))
void ref_make_groups(void* img_u16, int dim_x, int dim_y, void* coords_i32, void* g_info_f32, float pct_thresh) {
    CppGrouping g;
    g.make_groups(img_u16, dim_x, dim_y, coords_i32, g_info_f32, pct_thresh);
}
