// Validation metrics on the device (reference: utils.py:32-72 roc_auc_cuda / accuracy, called
// once per epoch from main.py:189-195,246-252 on ~1e4..1.5e6 predictions after a D2H copy + sklearn on the host).
//   AUROC  = area under sklearn.metrics.roc_curve (one point per DISTINCT score, (0, 0) prepended, trapezoids)
//   AUPR   = sklearn.metrics.average_precision_score = sum_k (R_k - R_{k-1}) P_k over distinct-score thresholds
//   accuracy = mean[(pred >= 0.5) == (label >= 0.5)]
// each over all samples and per hyperedge size.  Pipeline, all hand-written kernels on the caller's stream:
//   key = (class << 32) | ~orderable(score)  ->  stable LSD radix sort (8-bit digits: per-tile histograms, one scan,
//   warp-ordered stable scatter)  ->  one packed scan (positives so far | tie groups so far)  ->  one thread per tie group
//   adds its trapezoid / precision step in fp64.  Ties are exact (equal fp32 scores share one threshold, -0.0 == +0.0).
#include "common.cuh"

namespace matcha {
namespace {

constexpr int kTile = 2048;          // elements per block of a sort pass
constexpr int kSortThreads = 256;    // 8 warps x 8 rounds of 32 consecutive elements
constexpr int kMaxClasses = 64;

__device__ __forceinline__ uint32_t desc_key(float s) {
  uint32_t u = __float_as_uint(s);
  if (u == 0x80000000u) u = 0u;                                   // -0.0 and +0.0 are one threshold
  const uint32_t asc = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ~asc;                                                     // ascending key order = descending score
}

__global__ void metrics_keys_kernel(const float* __restrict__ score, const float* __restrict__ label, const int32_t* __restrict__ cls,
                                    int64_t n, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t c = cls ? (uint64_t)(uint32_t)cls[i] : 0ull;
    keys[i] = (c << 32) | desc_key(score[i]);
    vals[i] = label[i] > 0.5f ? 1u : 0u;
  }
}

// ---------------- stable LSD radix sort pass (8-bit digit at `shift`) ----------------
__global__ void __launch_bounds__(kSortThreads) sort_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift,
                                                                 uint32_t* __restrict__ hist, int nblocks) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kTile;
  for (int i = threadIdx.x; i < kTile; i += kSortThreads) {
    const int64_t idx = base + i;
    if (idx < n) atomicAdd(&h[(uint32_t)(keys[idx] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];      // digit-major: the scan order of an LSD pass
}
// exclusive scan of a uint32 array in place, one block (len = 256 * nblocks: a few 1e5 entries at most)
__global__ void __launch_bounds__(1024) scan_u32_kernel(uint32_t* __restrict__ a, int64_t len) {
  __shared__ uint32_t part[1024];
  const int64_t per = (len + 1023) / 1024;
  const int64_t b = (int64_t)threadIdx.x * per, e = (b + per < len) ? b + per : len;
  uint32_t s = 0;
  for (int64_t i = b; i < e; ++i) s += a[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {            // Hillis-Steele inclusive scan of the 1024 partials
    const uint32_t v = threadIdx.x >= off ? part[threadIdx.x - off] : 0u;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = threadIdx.x ? part[threadIdx.x - 1] : 0u;
  for (int64_t i = b; i < e; ++i) { const uint32_t v = a[i]; a[i] = run; run += v; }
}
__global__ void __launch_bounds__(kSortThreads) sort_scatter_kernel(const uint64_t* __restrict__ kin, const uint32_t* __restrict__ vin,
                                                                    uint64_t* __restrict__ kout, uint32_t* __restrict__ vout, int64_t n,
                                                                    int shift, const uint32_t* __restrict__ hist, int nblocks) {
  __shared__ uint32_t wh[kSortThreads / 32][256];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (kSortThreads / 32) * 256; i += kSortThreads) (&wh[0][0])[i] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kTile + w * (kTile / (kSortThreads / 32));
  constexpr int kRounds = kTile / kSortThreads;
  uint64_t key[kRounds];
  uint32_t val[kRounds];
  // phase 1: each warp counts the digits of its contiguous sub-tile
#pragma unroll
  for (int r = 0; r < kRounds; ++r) {
    const int64_t idx = base + r * 32 + lane;
    const bool valid = idx < n;
    key[r] = valid ? kin[idx] : 0ull;
    val[r] = valid ? vin[idx] : 0u;
    const uint32_t d = valid ? ((uint32_t)(key[r] >> shift) & 255u) : 256u + lane;      // invalid lanes match nobody
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    if (valid && lane == __ffs(peers) - 1) wh[w][d] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // phase 2: digit d of warp w starts at (global start of d for this block) + (counts of d in the warps before w)
  {
    const int d = threadIdx.x;
    uint32_t run = hist[(int64_t)d * nblocks + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < kSortThreads / 32; ++ww) { const uint32_t c = wh[ww][d]; wh[ww][d] = run; run += c; }
  }
  __syncthreads();
  // phase 3: replay in the same order; rank among equal digits of a round = position among the matching lanes
#pragma unroll
  for (int r = 0; r < kRounds; ++r) {
    const int64_t idx = base + r * 32 + lane;
    const bool valid = idx < n;
    const uint32_t d = valid ? ((uint32_t)(key[r] >> shift) & 255u) : 256u + lane;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    uint32_t pos = 0;
    if (valid) pos = wh[w][d] + __popc(peers & ((1u << lane) - 1u));
    __syncwarp();
    if (valid && lane == __ffs(peers) - 1) wh[w][d] += __popc(peers);
    __syncwarp();
    if (valid) { kout[pos] = key[r]; vout[pos] = val[r]; }
  }
}

// ---------------- packed scan over the sorted array: (positives so far) << 32 | (tie-group ends so far) ----------------
__device__ __forceinline__ uint64_t scan_item(const uint64_t* keys, const uint32_t* vals, int64_t i, int64_t n) {
  const uint64_t end = (i == n - 1 || keys[i] != keys[i + 1]) ? 1ull : 0ull;
  return ((uint64_t)vals[i] << 32) | end;
}
__global__ void __launch_bounds__(256) scan_partial_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                           int64_t n, uint64_t* __restrict__ block_sum) {
  __shared__ uint64_t red[256];
  const int64_t base = (int64_t)blockIdx.x * kTile;
  uint64_t s = 0;
  for (int i = threadIdx.x; i < kTile; i += 256) { const int64_t idx = base + i; if (idx < n) s += scan_item(keys, vals, idx, n); }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) { if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off]; __syncthreads(); }
  if (threadIdx.x == 0) block_sum[blockIdx.x] = red[0];
}
__global__ void __launch_bounds__(1024) scan_u64_kernel(uint64_t* __restrict__ a, int64_t len) {      // exclusive, in place, one block
  __shared__ uint64_t part[1024];
  const int64_t per = (len + 1023) / 1024;
  const int64_t b = (int64_t)threadIdx.x * per, e = (b + per < len) ? b + per : len;
  uint64_t s = 0;
  for (int64_t i = b; i < e; ++i) s += a[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const uint64_t v = threadIdx.x >= off ? part[threadIdx.x - off] : 0ull;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  uint64_t run = threadIdx.x ? part[threadIdx.x - 1] : 0ull;
  for (int64_t i = b; i < e; ++i) { const uint64_t v = a[i]; a[i] = run; run += v; }
}
// inclusive packed scan within each tile (thread = 8 consecutive elements) + compaction of the tie-group ends:
//   gend[g] = sorted index of the end of group g,  gtp[g] = positives among sorted[0 .. gend[g]]
__global__ void __launch_bounds__(256) scan_apply_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, int64_t n,
                                                         const uint64_t* __restrict__ block_off, uint32_t* __restrict__ cum_tp,
                                                         uint32_t* __restrict__ gend, uint32_t* __restrict__ gtp) {
  __shared__ uint64_t part[256];
  constexpr int kPer = kTile / 256;
  const int64_t base = (int64_t)blockIdx.x * kTile + threadIdx.x * kPer;
  uint64_t it[kPer];
  uint64_t s = 0;
#pragma unroll
  for (int j = 0; j < kPer; ++j) { it[j] = (base + j < n) ? scan_item(keys, vals, base + j, n) : 0ull; s += it[j]; }
  part[threadIdx.x] = s;
  __syncthreads();
  for (int off = 1; off < 256; off <<= 1) {
    const uint64_t v = threadIdx.x >= off ? part[threadIdx.x - off] : 0ull;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  uint64_t run = block_off[blockIdx.x] + (threadIdx.x ? part[threadIdx.x - 1] : 0ull);
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    if (base + j >= n) break;
    const uint32_t g_before = (uint32_t)run;
    run += it[j];
    const uint32_t tp = (uint32_t)(run >> 32);
    cum_tp[base + j] = tp;
    if (it[j] & 1ull) { gend[g_before] = (uint32_t)(base + j); gtp[g_before] = tp; }
  }
}

// per class: first sorted index, positives before it, count, positives  (binary search on the class bits)
struct ClassSeg { uint32_t start, count, base_tp, pos; };
__global__ void class_seg_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ cum_tp, int64_t n, int n_classes,
                                 ClassSeg* __restrict__ seg) {
  const int c = threadIdx.x;
  if (c >= n_classes) return;
  auto lower = [&](uint64_t k) { int64_t lo = 0, hi = n; while (lo < hi) { const int64_t m = (lo + hi) >> 1; if (keys[m] < k) lo = m + 1; else hi = m; } return lo; };
  const int64_t s = lower((uint64_t)c << 32), e = lower((uint64_t)(c + 1) << 32);
  ClassSeg g;
  g.start = (uint32_t)s; g.count = (uint32_t)(e - s);
  g.base_tp = s > 0 ? cum_tp[s - 1] : 0u;
  g.pos = e > s ? cum_tp[e - 1] - g.base_tp : 0u;
  seg[c] = g;
}
// one thread per tie group: trapezoid of the ROC curve and precision step of the PR curve, fp64, per class
__global__ void __launch_bounds__(256) group_accum_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ gend,
                                                          const uint32_t* __restrict__ gtp, const uint32_t* __restrict__ cum_tp, int64_t n,
                                                          const ClassSeg* __restrict__ seg, int n_classes, double* __restrict__ acc /*[n_classes][2]*/) {
  __shared__ double sh[kMaxClasses][2];
  for (int i = threadIdx.x; i < kMaxClasses * 2; i += 256) (&sh[0][0])[i] = 0.0;
  __syncthreads();
  const int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x;
  // exact number of tie groups: left in seg[n_classes].start by store_group_count_kernel (the grid is sized for n)
  const uint32_t n_groups = seg[n_classes].start;
  if (g < n_groups) {
    const uint32_t i = gend[g];
    const int c = (int)(keys[i] >> 32);
    if (c < n_classes) {
      const ClassSeg s = seg[c];
      const uint32_t P = s.pos, N = s.count - s.pos;
      uint32_t tp_prev = 0, r_prev = 0;
      if (g > 0) {
        const uint32_t ip = gend[g - 1];
        if (ip >= s.start) { tp_prev = gtp[g - 1] - s.base_tp; r_prev = ip - s.start + 1; }
      }
      const uint32_t tp = gtp[g] - s.base_tp, r = i - s.start + 1;
      const uint32_t fp = r - tp, fp_prev = r_prev - tp_prev;
      if (P > 0 && N > 0) {
        atomicAdd(&sh[c][0], (double)(fp - fp_prev) * ((double)tp + (double)tp_prev) * 0.5 / ((double)P * (double)N));
        atomicAdd(&sh[c][1], ((double)(tp - tp_prev) / (double)P) * ((double)tp / (double)r));
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_classes * 2; i += 256) {
    const double v = (&sh[0][0])[i];
    if (v != 0.0) atomicAdd(&acc[i], v);
  }
}
__global__ void store_group_count_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, int64_t n,
                                         const uint64_t* __restrict__ block_off, int nblocks, const uint64_t* __restrict__ last_block_sum,
                                         ClassSeg* __restrict__ seg, int n_classes) {
  // total tie groups = exclusive offset of the last tile + its own sum
  if (threadIdx.x == 0 && blockIdx.x == 0) seg[n_classes].start = (uint32_t)(block_off[nblocks - 1] + last_block_sum[0]);
}
__global__ void __launch_bounds__(256) accuracy_kernel(const float* __restrict__ score, const float* __restrict__ label,
                                                       const int32_t* __restrict__ cls, int64_t n, int n_classes,
                                                       unsigned long long* __restrict__ cnt /*[n_classes + 1][2]: correct, count*/) {
  __shared__ unsigned int sh[kMaxClasses + 1][2];
  for (int i = threadIdx.x; i < (kMaxClasses + 1) * 2; i += 256) (&sh[0][0])[i] = 0u;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const bool ok = (score[i] >= 0.5f) == (label[i] >= 0.5f);
    atomicAdd(&sh[0][1], 1u);
    if (ok) atomicAdd(&sh[0][0], 1u);
    if (cls) {
      const int c = cls[i];
      if (c >= 0 && c < n_classes) { atomicAdd(&sh[1 + c][1], 1u); if (ok) atomicAdd(&sh[1 + c][0], 1u); }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (n_classes + 1) * 2; i += 256) {
    const unsigned int v = (&sh[0][0])[i];
    if (v) atomicAdd(&cnt[i], (unsigned long long)v);
  }
}
// out rows: 0 = all, 1 + c = class c; columns {auroc, aupr, accuracy, count}
__global__ void finalize_kernel(const double* __restrict__ acc_all, const ClassSeg* __restrict__ seg_all, const double* __restrict__ acc_cls,
                                const ClassSeg* __restrict__ seg_cls, const unsigned long long* __restrict__ cnt, int n_classes,
                                double* __restrict__ out) {
  const int r = threadIdx.x;
  if (r > n_classes) return;
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  const ClassSeg s = r == 0 ? seg_all[0] : (seg_cls ? seg_cls[r - 1] : ClassSeg{0u, 0u, 0u, 0u});
  const double* a = r == 0 ? acc_all : (acc_cls ? acc_cls + 2 * (r - 1) : nullptr);
  const bool both = s.pos > 0 && s.count > s.pos;
  out[4 * r + 0] = (a && both) ? a[0] : nan;
  out[4 * r + 1] = (a && both) ? a[1] : nan;
  const unsigned long long c = cnt[2 * r + 1];
  out[4 * r + 2] = c ? (double)cnt[2 * r] / (double)c : nan;
  out[4 * r + 3] = (double)c;
}

struct MetricsWs {
  uint64_t *keysA, *keysB, *block_sum, *block_sum_keep;
  uint32_t *valsA, *valsB, *hist, *cum_tp, *gend, *gtp;
  ClassSeg *seg_all, *seg_cls;
  double *acc_all, *acc_cls;
  unsigned long long* cnt;
  int64_t bytes;
};
MetricsWs carve_metrics(int64_t n, void* base) {
  MetricsWs w;
  char* p = reinterpret_cast<char*>(base);
  int64_t off = 0;
  auto take = [&](int64_t nbytes) -> void* { void* r = p ? (void*)(p + off) : nullptr; off += (nbytes + 255) / 256 * 256; return r; };
  const int64_t nb = (n + kTile - 1) / kTile;
  w.keysA = (uint64_t*)take(8 * n); w.keysB = (uint64_t*)take(8 * n);
  w.valsA = (uint32_t*)take(4 * n); w.valsB = (uint32_t*)take(4 * n);
  w.hist = (uint32_t*)take(4 * 256 * nb);
  w.block_sum = (uint64_t*)take(8 * (nb + 1)); w.block_sum_keep = (uint64_t*)take(8 * (nb + 1));
  w.cum_tp = (uint32_t*)take(4 * n); w.gend = (uint32_t*)take(4 * n); w.gtp = (uint32_t*)take(4 * n);
  w.seg_all = (ClassSeg*)take(sizeof(ClassSeg) * 2); w.seg_cls = (ClassSeg*)take(sizeof(ClassSeg) * (kMaxClasses + 1));
  w.acc_all = (double*)take(8 * 2); w.acc_cls = (double*)take(8 * 2 * kMaxClasses);
  w.cnt = (unsigned long long*)take(8 * 2 * (kMaxClasses + 1));
  w.bytes = off;
  return w;
}

// sort (keysA, valsA) by bits [0, 8 * passes); result ends in (*ko, *vo)
int radix_sort(MetricsWs& w, int64_t n, int passes, uint64_t** ko, uint32_t** vo, cudaStream_t s) {
  const int nb = (int)((n + kTile - 1) / kTile);
  uint64_t *kin = w.keysA, *kout = w.keysB;
  uint32_t *vin = w.valsA, *vout = w.valsB;
  for (int p = 0; p < passes; ++p) {
    sort_hist_kernel<<<nb, kSortThreads, 0, s>>>(kin, n, 8 * p, w.hist, nb);
    MATCHA_CHECK_LAUNCH("sort_hist");
    scan_u32_kernel<<<1, 1024, 0, s>>>(w.hist, (int64_t)256 * nb);
    MATCHA_CHECK_LAUNCH("sort_scan");
    sort_scatter_kernel<<<nb, kSortThreads, 0, s>>>(kin, vin, kout, vout, n, 8 * p, w.hist, nb);
    MATCHA_CHECK_LAUNCH("sort_scatter");
    uint64_t* tk = kin; kin = kout; kout = tk;
    uint32_t* tv = vin; vin = vout; vout = tv;
  }
  *ko = kin; *vo = vin;
  return MATCHA_OK;
}

// curves of one keying (cls == nullptr: everything is class 0)
int run_curves(MetricsWs& w, const float* score, const float* label, const int32_t* cls, int64_t n, int n_classes, ClassSeg* seg,
               double* acc, cudaStream_t s) {
  int blocks = (int)((n + 255) / 256);
  if (blocks > kSMs * 8) blocks = kSMs * 8;
  metrics_keys_kernel<<<blocks, 256, 0, s>>>(score, label, cls, n, w.keysA, w.valsA);
  MATCHA_CHECK_LAUNCH("metrics_keys");
  uint64_t* k; uint32_t* v;
  if (int rc = radix_sort(w, n, cls ? 5 : 4, &k, &v, s)) return rc;
  const int nb = (int)((n + kTile - 1) / kTile);
  scan_partial_kernel<<<nb, 256, 0, s>>>(k, v, n, w.block_sum);
  MATCHA_CHECK_LAUNCH("scan_partial");
  if (int rc = check_cuda(cudaMemcpyAsync(w.block_sum_keep, w.block_sum + (nb - 1), 8, cudaMemcpyDeviceToDevice, s), "copy last tile sum")) return rc;
  scan_u64_kernel<<<1, 1024, 0, s>>>(w.block_sum, nb);
  MATCHA_CHECK_LAUNCH("scan_u64");
  scan_apply_kernel<<<nb, 256, 0, s>>>(k, v, n, w.block_sum, w.cum_tp, w.gend, w.gtp);
  MATCHA_CHECK_LAUNCH("scan_apply");
  class_seg_kernel<<<1, kMaxClasses, 0, s>>>(k, w.cum_tp, n, n_classes, seg);
  MATCHA_CHECK_LAUNCH("class_seg");
  store_group_count_kernel<<<1, 32, 0, s>>>(k, v, n, w.block_sum, nb, w.block_sum_keep, seg, n_classes);
  MATCHA_CHECK_LAUNCH("store_group_count");
  if (int rc = check_cuda(cudaMemsetAsync(acc, 0, sizeof(double) * 2 * n_classes, s), "memset metric accumulators")) return rc;
  group_accum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(k, w.gend, w.gtp, w.cum_tp, n, seg, n_classes, acc);
  MATCHA_CHECK_LAUNCH("group_accum");
  return MATCHA_OK;
}

}  // namespace
}  // namespace matcha

using namespace matcha;

extern "C" {

int64_t matcha_metrics_workspace_bytes(int64_t n) { return n < 0 ? -1 : carve_metrics(n > 0 ? n : 1, nullptr).bytes + 256; }

int matcha_binary_metrics(const float* score, const float* label, const int32_t* cls, int64_t n, int32_t n_classes, double* out,
                          void* workspace, int64_t workspace_bytes, void* stream) {
  MATCHA_REQUIRE(score && label && out && workspace && n > 0, "matcha_binary_metrics: bad arguments");
  MATCHA_REQUIRE(n < (1ll << 31), "matcha_binary_metrics: n = %lld exceeds 2^31 - 1", (long long)n);
  MATCHA_REQUIRE(n_classes >= 0 && n_classes <= kMaxClasses && (cls || n_classes == 0), "matcha_binary_metrics: n_classes=%d out of range (0..%d)",
                 (int)n_classes, kMaxClasses);
  cudaStream_t s = (cudaStream_t)stream;
  uintptr_t base = ((uintptr_t)workspace + 255) / 256 * 256;
  MetricsWs w = carve_metrics(n, (void*)base);
  MATCHA_REQUIRE((int64_t)(base - (uintptr_t)workspace) + w.bytes <= workspace_bytes, "matcha_binary_metrics: workspace too small: need %lld bytes",
                 (long long)(w.bytes + 256));
  int rc;
  if ((rc = run_curves(w, score, label, nullptr, n, 1, w.seg_all, w.acc_all, s))) return rc;
  if (n_classes > 0 && (rc = run_curves(w, score, label, cls, n, n_classes, w.seg_cls, w.acc_cls, s))) return rc;
  if ((rc = check_cuda(cudaMemsetAsync(w.cnt, 0, 8 * 2 * (kMaxClasses + 1), s), "memset accuracy counters"))) return rc;
  int blocks = (int)((n + 255) / 256);
  if (blocks > kSMs * 4) blocks = kSMs * 4;
  accuracy_kernel<<<blocks, 256, 0, s>>>(score, label, n_classes > 0 ? cls : nullptr, n, n_classes, w.cnt);
  MATCHA_CHECK_LAUNCH("accuracy");
  finalize_kernel<<<1, kMaxClasses + 1, 0, s>>>(w.acc_all, w.seg_all, n_classes > 0 ? w.acc_cls : nullptr, n_classes > 0 ? w.seg_cls : nullptr,
                                                w.cnt, n_classes, out);
  MATCHA_CHECK_LAUNCH("metrics_finalize");
  return MATCHA_OK;
}

}  // extern "C"
