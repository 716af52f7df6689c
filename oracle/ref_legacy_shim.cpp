// ref_legacy_shim.cpp -- extern "C" doorway into the reference's legacy (un-imported) twins
// under cpp_wrappers/, compiled where they lie by oracle/Makefile into
// oracle/_ref/libref_legacy.so.  TEST INFRASTRUCTURE ONLY.
//   batch_ordered_neighbors (cpp_wrappers/cpp_neighbors/neighbors/neighbors.cpp:125-208):
//     brute-force radius search whose tie order is deterministic (ascending support index);
//     the cross-check for the tie-break the CUDA kernel and oracle/lcr_oracle.c use.
//   grid_subsampling        (cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-106)
// A separate library because cpp_wrappers/cpp_utils/cloud/cloud.h defines its own PointXYZ.
#include <cstdint>
#include <cstring>
#include <vector>

#include "cpp_neighbors/neighbors/neighbors.h"
#include "cpp_subsampling/grid_subsampling/grid_subsampling.h"

extern "C" {

static std::vector<int> g_last;
int64_t ref_batch_ordered_neighbors(const float* q, int64_t nq, const float* s, int64_t ns,
                                    const int64_t* q_len, const int64_t* s_len, int batch,
                                    float radius, int32_t* out_idx) {
  if (out_idx == nullptr) {
    std::vector<PointXYZ> vq(reinterpret_cast<const PointXYZ*>(q),
                             reinterpret_cast<const PointXYZ*>(q) + nq);
    std::vector<PointXYZ> vs(reinterpret_cast<const PointXYZ*>(s),
                             reinterpret_cast<const PointXYZ*>(s) + ns);
    std::vector<int> ql(q_len, q_len + batch), sl(s_len, s_len + batch);
    g_last.clear();
    batch_ordered_neighbors(vq, vs, ql, sl, g_last, radius);
    return nq > 0 ? static_cast<int64_t>(g_last.size() / nq) : 0;
  }
  std::memcpy(out_idx, g_last.data(), sizeof(int) * g_last.size());
  return nq > 0 ? static_cast<int64_t>(g_last.size() / nq) : 0;
}

// Legacy single-cloud grid subsampling without features/labels (sampleDl, verbose=0).
int64_t ref_legacy_grid_subsampling(const float* pts, int64_t n, float voxel, float* out_pts) {
  std::vector<PointXYZ> v(reinterpret_cast<const PointXYZ*>(pts),
                          reinterpret_cast<const PointXYZ*>(pts) + n);
  std::vector<PointXYZ> s;
  std::vector<float> f, sf;
  std::vector<int> c, sc;
  grid_subsampling(v, s, f, sf, c, sc, voxel, 0);
  std::memcpy(out_pts, s.data(), sizeof(float) * 3 * s.size());
  return static_cast<int64_t>(s.size());
}
}
