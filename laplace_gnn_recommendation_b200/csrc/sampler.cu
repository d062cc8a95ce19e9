// Sub-graph batch assembly on the device: the two index-heavy steps of the reference's per-user Python
// sampler, data/dataset.py:39-182,258-300 (GraphDataset.__getitem__, fetch_n_hop_neighbourhood, remap_edges_to_start_from_zero),
// for a whole batch of root users at once:
//   lgb_segment_expand      neighbour lists of a frontier: for every frontier node i (with its CSR row nodes[i]) emit one
//                           (i, neighbour) pair per adjacency entry, in adjacency order -- create_neighbouring_article_edges +
//                           the flatten() of fetch_n_hop_neighbourhood for all users of all roots in one launch;
//   lgb_bucketize_segmented remap_edges_to_start_from_zero for a batch: t.bucketize(value, buckets_of_its_root) where the
//                           sorted unique node ids of every root are concatenated in one array.
// Both are pure index arithmetic (bit-exact), one thread per output element, coalesced writes.
#include "common.cuh"

namespace lgb {

// out_off is the exclusive scan of the frontier nodes' degrees ([n + 1]); element j belongs to the node i with
// out_off[i] <= j < out_off[i + 1].
__global__ void __launch_bounds__(256) segment_expand_kernel(const int64_t* __restrict__ ptr, const int64_t* __restrict__ idx,
                                                             const int64_t* __restrict__ nodes, const int64_t* __restrict__ out_off,
                                                             int64_t n, int64_t total, int64_t* __restrict__ out_pos,
                                                             int64_t* __restrict__ out_nbr) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= total) return;
  int64_t lo = 0, hi = n;          // invariant: out_off[lo] <= j < out_off[hi]
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (out_off[mid] <= j) lo = mid; else hi = mid;
  }
  out_pos[j] = lo;
  out_nbr[j] = idx[ptr[nodes[lo]] + (j - out_off[lo])];
}

// out[j] = number of buckets of segment seg[j] that are < values[j]  (torch.bucketize(v, boundaries, right=False))
__global__ void __launch_bounds__(256) bucketize_segmented_kernel(const int64_t* __restrict__ values, const int64_t* __restrict__ seg,
                                                                  int64_t n, const int64_t* __restrict__ buckets,
                                                                  const int64_t* __restrict__ bucket_ptr, int64_t* __restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int64_t v = values[j];
  const int64_t b0 = bucket_ptr[seg[j]];
  int64_t lo = b0, hi = bucket_ptr[seg[j] + 1];
  while (lo < hi) {                // lower bound
    const int64_t mid = (lo + hi) >> 1;
    if (buckets[mid] < v) lo = mid + 1; else hi = mid;
  }
  out[j] = lo - b0;
}

}  // namespace lgb

using namespace lgb;

extern "C" {

int lgb_segment_expand(const int64_t* ptr, const int64_t* idx, const int64_t* nodes, const int64_t* out_off, int64_t n,
                       int64_t total, int64_t* out_pos, int64_t* out_nbr, void* stream) {
  LGB_REQUIRE(n >= 0 && total >= 0 && (total == 0 || (ptr && idx && nodes && out_off && out_pos && out_nbr)), LGB_EINVAL,
              "lgb_segment_expand: bad argument");
  if (total == 0) return LGB_OK;
  segment_expand_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ptr, idx, nodes, out_off, n, total, out_pos,
                                                                                           out_nbr);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

int lgb_bucketize_segmented(const int64_t* values, const int64_t* seg, int64_t n, const int64_t* buckets, const int64_t* bucket_ptr,
                            int64_t* out, void* stream) {
  LGB_REQUIRE(n >= 0 && (n == 0 || (values && seg && buckets && bucket_ptr && out)), LGB_EINVAL, "lgb_bucketize_segmented: bad argument");
  if (n == 0) return LGB_OK;
  bucketize_segmented_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(values, seg, n, buckets, bucket_ptr, out);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

}  // extern "C"
