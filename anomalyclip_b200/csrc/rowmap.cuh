// Row order bookkeeping of the temporal path.
//
// The caller's rows are ordered "(b n s l)" (n segments, s sub-videos, l frames per segment);
// the reference regroups them to "(b s) n l" before the axial transformer and back afterwards
// (src/models/components/temporal_model.py:46-53,67-71).  The kernels keep everything in
// sub-video order, so the regrouping happens once on the way in and once on the way out.
#pragma once

namespace aclip {

struct RowMap {
  int n, s, l;      // num_segments, segment_size, seg_length
  long long row0;   // first sub-video-order row of the chunk being processed

  // sub-video-order row  ((b*s + j)*n + i)*l + k   ->   caller row ((b*n + i)*s + j)*l + k
  // r is chunk-local
  __host__ __device__ long long caller_row(long long r) const {
    r += row0;
    const long long k = r % l;
    long long t = r / l;
    const long long i = t % n;
    t /= n;
    const long long j = t % s;
    const long long b = t / s;
    return ((b * n + i) * s + j) * l + k;
  }
};

}  // namespace aclip
