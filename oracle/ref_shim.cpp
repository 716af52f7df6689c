// ref_shim.cpp -- extern "C" doorway into the UNMODIFIED reference sources, compiled where
// they lie under /root/reference by oracle/Makefile into oracle/_ref/libref_ext.so.
// TEST INFRASTRUCTURE ONLY (validates oracle/lcr_oracle.c, optional CPU baseline).
// No reference source is copied into this repository; this file only declares and calls
//   grid_subsampling_cpu   (utils/extensions/cpu/grid_subsampling/grid_subsampling_cpu.h:30-36)
//   radius_neighbors_cpu   (utils/extensions/cpu/radius_neighbors/radius_neighbors_cpu.h:9-16)
// i.e. exactly what utils/extensions/cpu/*/{grid_subsampling,radius_neighbors}.cpp call
// after unpacking their at::Tensor arguments (grid_subsampling.cpp:20-38,
// radius_neighbors.cpp:29-55).
#include <cstdint>
#include <cstring>
#include <vector>

#include "cpu/grid_subsampling/grid_subsampling_cpu.h"
#include "cpu/radius_neighbors/radius_neighbors_cpu.h"

extern "C" {

// Returns total output points; out_pts must hold 3*n_total floats.
int64_t ref_grid_subsampling(const float* pts, int64_t n_total, const int64_t* lengths,
                             int batch, float voxel, float* out_pts, int64_t* out_lengths) {
  std::vector<PointXYZ> v(reinterpret_cast<const PointXYZ*>(pts),
                          reinterpret_cast<const PointXYZ*>(pts) + n_total);
  std::vector<PointXYZ> s;
  std::vector<long> len(lengths, lengths + batch), slen;
  grid_subsampling_cpu(v, s, len, slen, voxel);
  std::memcpy(out_pts, s.data(), sizeof(float) * 3 * s.size());
  for (int b = 0; b < batch; b++) out_lengths[b] = slen[b];
  return static_cast<int64_t>(s.size());
}

// Two-call protocol: first call with out_idx == nullptr returns max_count and keeps the
// table in a static buffer; second call copies it out ([nq, max_count] int64).
static std::vector<long> g_last;
int64_t ref_radius_neighbors(const float* q, int64_t nq, const float* s, int64_t ns,
                             const int64_t* q_len, const int64_t* s_len, int batch,
                             float radius, int64_t* out_idx) {
  if (out_idx == nullptr) {
    std::vector<PointXYZ> vq(reinterpret_cast<const PointXYZ*>(q),
                             reinterpret_cast<const PointXYZ*>(q) + nq);
    std::vector<PointXYZ> vs(reinterpret_cast<const PointXYZ*>(s),
                             reinterpret_cast<const PointXYZ*>(s) + ns);
    std::vector<long> ql(q_len, q_len + batch), sl(s_len, s_len + batch);
    g_last.clear();
    radius_neighbors_cpu(vq, vs, ql, sl, g_last, radius);
    return nq > 0 ? static_cast<int64_t>(g_last.size() / nq) : 0;
  }
  std::memcpy(out_idx, g_last.data(), sizeof(long) * g_last.size());
  return nq > 0 ? static_cast<int64_t>(g_last.size() / nq) : 0;
}

}  // extern "C"
